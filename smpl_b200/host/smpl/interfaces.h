// Our own copies of the reference's three plugin interfaces, signature for signature, so that the
// adapters in gpu_adapters.h are drop-in implementations (SBPL, ROS and Eigen are not installed here,
// and the reference headers pull them in).  Only what the hot path touches is declared.
//
//   sbpl::motion::Extension                    smpl/include/smpl/extension.h:40-60
//   sbpl::motion::RobotState, GoalConstraint   smpl/include/smpl/types.h:67, 178-194
//   sbpl::motion::CollisionChecker             smpl/include/smpl/collision_checker.h:48-130
//   sbpl::motion::RobotModel,
//     ForwardKinematicsInterface               smpl/include/smpl/robot_model.h:50-110
//   sbpl::motion::RobotHeuristic               smpl/include/smpl/heuristic/robot_heuristic.h:53-100
//     (the SBPL `Heuristic` base and RobotPlanningSpaceObserver are not available; the virtuals the
//      planner calls on the hot path are kept with identical names and meaning)
#ifndef SMPLHOST_SMPL_INTERFACES_H
#define SMPLHOST_SMPL_INTERFACES_H

#include <cstddef>
#include <cstdint>
#include <limits>
#include <string>
#include <typeinfo>
#include <vector>

namespace sbpl {
namespace motion {

template <typename T>
size_t GetClassCode()
{
    return typeid(T).hash_code();
}

class Extension
{
public:
    virtual ~Extension() { }

    template <typename T>
    T* getExtension()
    {
        Extension* e = getExtension(GetClassCode<T>());
        return dynamic_cast<T*>(e);
    }

    virtual Extension* getExtension(size_t class_code) = 0;
};

typedef std::vector<double> RobotState;

enum GoalType { INVALID_GOAL_TYPE = -1, XYZ_GOAL, XYZ_RPY_GOAL, JOINT_STATE_GOAL, NUMBER_OF_GOAL_TYPES };

struct GoalConstraint
{
    RobotState angles;
    std::vector<double> angle_tolerances;
    std::vector<double> pose;
    double xyz_offset[3];
    double xyz_tolerance[3];
    double rpy_tolerance[3];
    std::vector<double> tgt_off_pose;
    int xyz[3];
    GoalType type;
};

class CollisionChecker : public virtual Extension
{
public:
    virtual ~CollisionChecker() { }
    virtual bool isStateValid(const RobotState& state, bool verbose = false) = 0;
    virtual bool isStateValid(const RobotState& state, double& distToObst, bool verbose = false) = 0;
    virtual bool isStateToStateValid(const RobotState& start, const RobotState& finish, bool verbose = false) = 0;
    virtual bool isStateToStateValid(const RobotState& angles0, const RobotState& angles1, double& distToObst,
                                     int& distToObstCells, bool verbose = false) = 0;
    virtual bool interpolatePath(const RobotState& start, const RobotState& finish, std::vector<RobotState>& path) = 0;
    // fork hooks, no-ops in the reference (collision_checker.h:106-124)
    virtual void setLastExpansionStep(int) { }
    virtual void markGridForExpandedState(const RobotState&, const RobotState&, int) { }
    virtual void resetCellsMarking(int) { }
    virtual void setClearanceThreshold(double) { }
};

/// smpl/collision_checker.h:132-144
class CollisionDistanceExtension : public virtual Extension
{
public:
    virtual ~CollisionDistanceExtension() { }
    virtual double distanceToCollision(const RobotState& state) = 0;
    virtual double distanceToCollision(const RobotState& start, const RobotState& finish) = 0;
};

class RobotModel : public virtual Extension
{
public:
    virtual ~RobotModel() { }
    virtual double minPosLimit(int jidx) const = 0;
    virtual double maxPosLimit(int jidx) const = 0;
    virtual bool hasPosLimit(int jidx) const = 0;
    virtual bool isContinuous(int jidx) const = 0;
    virtual double velLimit(int jidx) const = 0;
    virtual double accLimit(int jidx) const = 0;
    virtual bool checkJointLimits(const RobotState& state, bool verbose = false) = 0;
    size_t jointCount() const { return planning_joints_.size(); }
    size_t jointVariableCount() const { return planning_joints_.size(); }
    void setPlanningJoints(const std::vector<std::string>& joints) { planning_joints_ = joints; }
    const std::vector<std::string>& getPlanningJoints() const { return planning_joints_; }
protected:
    std::vector<std::string> planning_joints_;
};

class ForwardKinematicsInterface : public virtual RobotModel
{
public:
    virtual ~ForwardKinematicsInterface() { }
    virtual bool computeFK(const RobotState& state, const std::string& name, std::vector<double>& pose) = 0;
    virtual bool computePlanningLinkFK(const RobotState& state, std::vector<double>& pose) = 0;
};

class RobotHeuristic : public virtual Extension
{
public:
    static const int Infinity = std::numeric_limits<int16_t>::max();
    virtual ~RobotHeuristic() { }
    virtual double getMetricStartDistance(double x, double y, double z) = 0;
    virtual double getMetricGoalDistance(double x, double y, double z) = 0;
    virtual void updateGoal(const GoalConstraint& goal) = 0;
    virtual int GetGoalHeuristic(int state_id) = 0;
    virtual int GetStartHeuristic(int state_id) = 0;
    virtual int GetFromToHeuristic(int from_id, int to_id) = 0;
};

} // namespace motion
} // namespace sbpl

#endif
