// ORACLE (test infrastructure only -- never linked into the product path).
//
// CPU restatement of sbpl::collision::{SelfCollisionModelImpl, CollisionSpace,
// AttachedBodiesCollisionModel/State, CheckVoxelsCollisions} and the subset of
// MoveIt's AllowedCollisionMatrix the reference uses.
// parity unpinned (no reference golden vectors).
#ifndef ORACLE_COLLISION_SPACE_H
#define ORACLE_COLLISION_SPACE_H

#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "collision_model.h"
#include "distance_map.h"

namespace oracle {

/// MoveIt collision_detection::AllowedCollisionMatrix, only setEntry /
/// getEntry / hasEntry (self_collision_model.cpp:285-307, 366-379, 1136, 1254).
class AllowedCollisionMatrix
{
public:
    enum Type { NEVER = 0, ALWAYS = 1 };
    void setEntry(const std::string& a, const std::string& b, bool allowed);
    void setEntry(const std::string& name, bool allowed);
    bool hasEntry(const std::string& name) const { return m_entries.count(name) != 0; }
    bool getEntry(const std::string& a, const std::string& b, Type& type) const;
    void clear() { m_entries.clear(); }
private:
    std::map<std::string, std::map<std::string, Type>> m_entries;
};

/// attached_bodies_collision_model.cpp:70-141, 264-313: a body attached to a
/// robot link, modelled by equal-radius spheres (one per voxel of the body).
struct AttachedBody
{
    std::string id;
    int link_index;
    SphereModelTree spheres;
    std::vector<SphereState> states;
};

/// OccupancyGrid (smpl/include/smpl/occupancy_grid.h) with ref_counted=false:
/// a thin wrapper; kept only so call sites read like the reference's.
struct OccupancyGrid
{
    explicit OccupancyGrid(EuclidDistanceMap* df) : m_grid(df) { }
    double getSquaredDist(double x, double y, double z) const { return m_grid->getMetricSquaredDistance(x, y, z); }
    double getDistance(int x, int y, int z) const { return m_grid->getDistance(x, y, z); }
    /// OccupancyGrid::getDistanceFromPoint (occupancy_grid.h:228-231) -> getMetricDistance (distance_map.hpp:281-286)
    double getDistanceFromPoint(double x, double y, double z) const { return m_grid->getDistance(x, y, z); }
    void addPointsToField(const std::vector<Vec3>& p) { m_grid->addPointsToMap(p); }
    void removePointsFromField(const std::vector<Vec3>& p) { m_grid->removePointsFromMap(p); }
    EuclidDistanceMap* m_grid;
};

struct CheckStats
{
    long long df_lookups;       // CheckSphereCollision calls actually made (with early-out)
    long long sphere_pair_tests;
    CheckStats() : df_lookups(0), sphere_pair_tests(0) { }
};

/// Result of the exhaustive (no early-out) evaluation of one state, used for
/// parity accounting: verdict, L = DF lookups the reference semantics
/// requires, and the smallest distance of any decision to its flip point.
struct StateReport
{
    bool valid;
    int required_lookups;
    double min_cell_boundary_margin;   // metres to the nearest worldToGrid truncation boundary
    double min_sphere_pair_margin;     // | ||c2-c1|| - (r1+r2) | in metres over visited pairs
};

class CollisionSpace
{
public:
    /// collision_space.cpp:689-739
    bool init(OccupancyGrid* grid, const RobotDesc& desc, const std::string& group_name,
              const std::vector<std::string>& planning_joints, std::string* err = nullptr);

    void setAllowedCollisionMatrix(const AllowedCollisionMatrix& acm); // self_collision_model.cpp:386-392
    void setPadding(double padding) { m_padding = padding; }
    bool setJointPosition(const std::string& name, double position);  // collision_space.cpp:49-61
    void setWorldToModelTransform(const Affine3& t);

    /// attach a body modelled by spheres (centres in the link frame)
    bool attachSpheres(const std::string& id, const std::vector<Vec3>& centers, double radius, const std::string& link_name);
    bool detachBody(const std::string& id);

    /// collision_space.cpp:532-536
    bool isStateValid(const std::vector<double>& state);
    /// collision_space.cpp:538-581
    bool isStateToStateValid(const std::vector<double>& start, const std::vector<double>& finish, int* waypoint_count = nullptr);
    /// CollisionSpace::collisionDistance (collision_space.cpp:496-500) -> SelfCollisionModelImpl::collisionDistance
    /// (self_collision_model.cpp:503-531, 1386-1468, 1512-1642)
    double collisionDistance(const std::vector<double>& state);
    /// same verdict, every waypoint visited in index order without early-out
    bool isStateToStateValidExhaustive(const std::vector<double>& start, const std::vector<double>& finish,
                                       int* waypoint_count, int* required_lookups);
    /// waypoints of an edge (for test comparison against device interpolation)
    void edgeWaypoints(const std::vector<double>& start, const std::vector<double>& finish,
                       std::vector<std::vector<double>>& out);

    StateReport reportState(const std::vector<double>& state);

    /// positions of every sphere-tree node of the group (robot trees in group
    /// order then attached bodies), after FK of `state`
    void sphereCenters(const std::vector<double>& state, std::vector<Vec3>& out);

    const RobotCollisionModel& model() const { return m_rcm; }
    const RobotMotionCollisionModel& motionModel() const { return *m_rmcm; }
    RobotCollisionState& state() { return *m_rcs; }
    int groupIndex() const { return m_gidx; }
    const std::vector<int>& planningVariables() const { return m_planning_joint_to_collision_model_indices; }
    const std::vector<std::pair<int, int>>& checkedSpheresStates() const { return m_checked_spheres_states; }
    const std::vector<AttachedBody>& attachedBodies() const { return m_attached; }
    const std::vector<int>& groupAttachedBodies() const { return m_group_attached; }
    const std::vector<std::pair<int, int>>& checkedAttachedRobot() const { return m_checked_ab_robot; }
    const std::vector<std::pair<int, int>>& checkedAttachedAttached() const { return m_checked_ab_ab; }
    const AllowedCollisionMatrix& acm() const { return m_acm; }
    double padding() const { return m_padding; }
    CheckStats stats;

private:
    OccupancyGrid* m_grid;
    RobotCollisionModel m_rcm;
    std::unique_ptr<RobotMotionCollisionModel> m_rmcm;
    std::unique_ptr<RobotCollisionState> m_rcs;
    std::vector<double> m_joint_vars;
    std::vector<int> m_planning_joint_to_collision_model_indices;
    int m_gidx;
    int m_active_gidx; // SelfCollisionModelImpl::m_gidx, -1 until the first check
    std::vector<int> m_voxels_indices;
    AllowedCollisionMatrix m_acm;
    double m_padding;
    std::vector<AttachedBody> m_attached;
    std::vector<int> m_group_attached;
    std::vector<std::pair<int, int>> m_checked_spheres_states;
    std::vector<std::pair<int, int>> m_checked_ab_ab, m_checked_ab_robot;

    struct SphereRef { int kind; int ss; int s; }; // kind 0 = robot spheres state, 1 = attached body
    const SphereModel& sphereModel(const SphereRef& r) const;
    const Vec3& spherePos(const SphereRef& r) const;
    void updateSphere(const SphereRef& r);

    void initAllowedCollisionMatrix();
    void updateState(const double* vals);
    void prepareState();
    void updateGroup(int gidx);
    void updateVoxelsStates();
    void updateCheckedSpheresIndices();
    bool checkVoxelsCollisions(std::vector<SphereRef>& q, double& dist);
    bool checkSpheresStateCollision(int kindA, int ss1i, int kindB, int ss2i, double& dist);
    bool checkCollision(double& dist);
};

} // namespace oracle

#endif
