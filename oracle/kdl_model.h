// ORACLE (test infrastructure only -- never linked into the product path).
//
// CPU restatement of sbpl::motion::KDLRobotModel's forward kinematics and
// joint-limit check (sbpl_kdl_robot_model/src/kdl_robot_model.cpp:59-158,
// 191-235, 326-337, 350-423) plus ManipLattice::computePlanningFrameFK /
// getTargetOffsetPose (smpl/src/graph/manip_lattice.cpp:1358-1373, 2297-2312)
// and BfsHeuristic (smpl/src/heuristic/bfs_heuristic.cpp).
//
// The arithmetic of the FK lives in orocos_kdl + kdl_parser, third-party
// dependencies that are NOT vendored under /root/reference and whose version
// is unpinned (sbpl_kdl_robot_model/package.xml: bare <depend>orocos_kdl,
// ROS Indigo era => KDL 1.3.x).  The published algorithm is restated here:
//   kdl_parser toKdl(): Joint(name, F.p, F.M*axis, RotAxis|TransAxis),
//                       Segment(child, joint, F_parent_jnt)
//   Segment: f_tip = joint.pose(0).Inverse() * F_parent_jnt; pose(q) = joint.pose(q) * f_tip
//   Joint::pose(q): RotAxis -> Frame(Rotation::Rot2(axis, q), origin);
//                   TransAxis -> Frame(origin + axis*q); None -> Identity
//   ChainFkSolverPos_recursive::JntToCart(q, out, segmentNr): product of segments 0..segmentNr-1
//   Frame*Frame = Frame(M1*M2, M1*p2 + p1), sums left to right.
// Quirk kept (kdl_robot_model.cpp:411 with the map built at :152-154): the
// planning link's *segment index* is passed as segmentNr, so the pose returned
// is that of the planning link's parent segment tip.
// parity unpinned: the reference's test_kdl_robot_model.cpp only prints.
#ifndef ORACLE_KDL_MODEL_H
#define ORACLE_KDL_MODEL_H

#include <map>
#include <memory>
#include <string>
#include <vector>

#include "bfs3d.h"
#include "collision_space.h"
#include "omath.h"
#include "robot_desc.h"

namespace oracle {

struct KdlFrame
{
    double M[9]; // row-major rotation
    double p[3];
    static KdlFrame Identity();
};
KdlFrame operator*(const KdlFrame& a, const KdlFrame& b);

class KDLRobotModel
{
public:
    bool init(const RobotDesc& desc, const std::vector<std::string>& planning_joints,
              const std::string& chain_root_link, const std::string& chain_tip_link, std::string* err = nullptr);
    void setKinematicsToPlanningTransform(const KdlFrame& f) { m_T_kin_to_planning = f; }
    bool setPlanningLink(const std::string& name);

    bool checkJointLimits(const std::vector<double>& angles) const; // :326-337
    bool computePlanningLinkFK(const std::vector<double>& angles, std::vector<double>& pose) const; // :400-423

    std::vector<double> min_limits, max_limits;
    std::vector<bool> continuous;

    struct Segment
    {
        std::string name;
        int joint_kind; // 0 none, 1 rot, 2 trans
        double axis[3];
        double origin[3];
        KdlFrame f_tip;
        int q_index;    // index into the planning joint vector, -1 if none
    };
    std::vector<Segment> segments;
    int planning_segment_nr; // value passed as segmentNr (= index of the planning link's segment)

private:
    std::vector<std::string> m_planning_joints;
    KdlFrame m_T_kin_to_planning;
    double normalizeAngle(double a, double a_min, double a_max) const;
};

/// manip_lattice.cpp:2297-2312
std::vector<double> GetTargetOffsetPose(const std::vector<double>& tip_pose, const double xyz_offset[3]);

/// bfs_heuristic.{h,cpp}
class BfsHeuristic
{
public:
    static const int Infinity = 32767; // robot_heuristic.h:62 (INT16_MAX)

    BfsHeuristic(const EuclidDistanceMap* grid, double inflation_radius, int cost_per_cell);
    void syncGridAndBfs();                 // :331-353
    bool updateGoal(double x, double y, double z); // :83-101 (goal.tgt_off_pose xyz)
    int getGoalHeuristicAt(double x, double y, double z) const; // :148-163 after projectToPoint
    double getMetricGoalDistance(double x, double y, double z) const; // :127-138
    int getBfsCostToGoal(int x, int y, int z) const; // :355-366
    BFS_3D* bfs() { return m_bfs.get(); }
    int wall_count;
private:
    const EuclidDistanceMap* m_grid;
    double m_inflation_radius;
    int m_cost_per_cell;
    std::unique_ptr<BFS_3D> m_bfs;
};

} // namespace oracle

#endif
