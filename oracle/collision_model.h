// ORACLE (test infrastructure only -- never linked into the product path).
//
// CPU restatement of the reference's robot collision model, state, sphere
// trees and motion model (sbpl_collision_checking).  Each function cites the
// reference file:line it follows.  parity unpinned: the reference ships no
// golden outputs for this path.
#ifndef ORACLE_COLLISION_MODEL_H
#define ORACLE_COLLISION_MODEL_H

#include <map>
#include <string>
#include <utility>
#include <vector>

#include "omath.h"
#include "robot_desc.h"

namespace oracle {

enum JointType { FIXED = 0, REVOLUTE, PRISMATIC, CONTINUOUS, PLANAR, FLOATING };

// which joint transform function the reference selects
// (robot_collision_model.cpp:331-359, 382-407; transform_functions.h:95-258)
enum JointFn { FN_FIXED = 0, FN_REV_X, FN_REV_Y, FN_REV_Z, FN_REV_GENERIC, FN_PRISMATIC, FN_PLANAR, FN_FLOATING };

Affine3 ComputeJointTransform(JointFn fn, const Affine3& origin, const Vec3& axis, const double* jvals);

/// base_collision_models.h:48-73
struct SphereModel
{
    std::string name;
    Vec3 center;
    double radius;
    int priority;
    int left, right; // indices into the owning tree, -1 for leaves
    SphereModel() : radius(0.0), priority(0), left(-1), right(-1) { }
    bool isLeaf() const { return left == right; }
};

/// base_collision_models.h:77-170; base_collision_models.cpp:184-222, 337-444, 569-641
class SphereModelTree
{
public:
    void buildFrom(const std::vector<SphereConfig>& spheres);
    int root() const { return (int)nodes.size() - 1; }
    std::vector<SphereModel> nodes;
private:
    int buildRecursive(std::vector<const SphereConfig*>::iterator first,
                       std::vector<const SphereConfig*>::iterator last);
};

void ComputeOptimalBoundingSphere(const SphereModel& s1, const SphereModel& s2, Vec3& c, double& r);

struct SpheresModel
{
    int link_index;
    SphereModelTree spheres;
};

struct VoxelsModel
{
    int link_index;
    double voxel_res;
    std::vector<Vec3> voxels; // link frame
};

struct GroupModel
{
    std::string name;
    std::vector<int> link_indices;
};

/// robot_collision_model.{h,cpp}
class RobotCollisionModel
{
public:
    bool init(const RobotDesc& desc, std::string* err = nullptr);

    // robot model
    std::string name, model_frame;
    std::vector<std::string> jvar_names;
    std::vector<bool> jvar_continuous, jvar_has_position_bounds;
    std::vector<double> jvar_min_positions, jvar_max_positions;
    std::vector<int> jvar_joint_indices;
    std::map<std::string, int> jvar_name_to_index;

    std::vector<std::string> joint_names;
    std::vector<Affine3> joint_origins;
    std::vector<Vec3> joint_axes;
    std::vector<JointType> joint_types;
    std::vector<JointFn> joint_fns;
    std::vector<std::pair<int, int>> joint_var_indices;
    std::vector<int> joint_parent_links, joint_child_links;

    std::vector<std::string> link_names;
    std::vector<int> link_parent_joints;
    std::vector<std::vector<int>> link_children_joints;
    std::map<std::string, int> link_name_to_index;

    // collision model
    std::vector<SpheresModel> spheres_models;
    std::vector<VoxelsModel> voxels_models;
    std::vector<GroupModel> group_models;
    std::map<std::string, int> group_name_to_index;
    std::vector<int> link_spheres_models; // link -> spheres model index or -1
    std::vector<int> link_voxels_models;

    int jointCount() const { return (int)joint_names.size(); }
    int linkCount() const { return (int)link_names.size(); }
    int jointVarCount() const { return (int)jvar_names.size(); }
    bool hasLink(const std::string& n) const { return link_name_to_index.count(n) != 0; }
    int linkIndex(const std::string& n) const { return link_name_to_index.at(n); }
    bool hasSpheresModel(int lidx) const { return link_spheres_models[lidx] >= 0; }
    bool hasGroup(const std::string& n) const { return group_name_to_index.count(n) != 0; }
    int groupIndex(const std::string& n) const { return group_name_to_index.at(n); }

private:
    void addJoint(const JointDesc* j, const std::string& name, JointType type);
    bool expandGroups(const std::vector<GroupConfig>& groups, std::vector<GroupConfig>& expanded, std::string* err) const;
};

/// base_collision_states.h
struct SphereState
{
    Vec3 pos;
    int version;
    SphereState() : version(-1) { }
};

/// robot_collision_state.{h,cpp}: lazy, dirty-flag forward kinematics
class RobotCollisionState
{
public:
    explicit RobotCollisionState(const RobotCollisionModel* model);

    const RobotCollisionModel* model() const { return m_model; }
    bool setWorldToModelTransform(const Affine3& transform); // identity / planar+floating translation only
    bool setJointVarPosition(int vidx, double position);
    bool setJointVarPositions(const double* positions);
    const std::vector<double>& jointVarPositions() const { return m_jvar_positions; }

    bool updateLinkTransform(int lidx);
    const Affine3& linkTransform(int lidx) const { return m_link_transforms[lidx]; }
    bool updateSphereState(int ssidx, int sidx);
    const Vec3& spherePos(int ssidx, int sidx) const { return m_sphere_states[ssidx][sidx].pos; }

    bool voxelsStateDirty(int vsidx) const { return m_dirty_voxels_states[vsidx]; }
    bool updateVoxelsState(int vsidx);
    const std::vector<Vec3>& voxelsState(int vsidx) const { return m_voxels_states[vsidx]; }

    const std::vector<int>& groupSpheresStateIndices(int gidx) const { return m_group_spheres_indices[gidx]; }
    const std::vector<int>& groupOutsideVoxelsStateIndices(int gidx) const { return m_group_voxels_indices[gidx]; }

    // instrumentation for the CPU baseline / algorithmic byte counts
    long long link_transform_updates;

private:
    const RobotCollisionModel* m_model;
    std::vector<double> m_jvar_positions;
    std::vector<bool> m_dirty_link_transforms;
    std::vector<Affine3> m_link_transforms;
    std::vector<int> m_link_transform_versions;
    std::vector<bool> m_dirty_joint_transforms;
    std::vector<Affine3> m_joint_transforms;
    std::vector<std::vector<SphereState>> m_sphere_states;
    std::vector<bool> m_dirty_voxels_states;
    std::vector<std::vector<Vec3>> m_voxels_states;
    std::vector<int> m_link_voxels_states;
    std::vector<std::vector<int>> m_group_spheres_indices, m_group_voxels_indices;
    std::vector<int> m_q;
};

/// robot_motion_collision_model.{h,cpp}
class RobotMotionCollisionModel
{
public:
    explicit RobotMotionCollisionModel(const RobotCollisionModel* rcm);
    double getMaxSphereMotion(const std::vector<double>& start, const std::vector<double>& finish,
                              const std::vector<int>& variables) const;
    std::vector<Vec3> mr_centers;
    std::vector<double> mr_radii;
    std::vector<Vec3> m_centers;
    std::vector<double> m_radii;
private:
    const RobotCollisionModel* m_rcm;
};

/// robot_motion_collision_model.h:60-150 (MotionInterpolation, planning-variable subset overloads)
class MotionInterpolation
{
public:
    explicit MotionInterpolation(const RobotCollisionModel* rcm) : m_rcm(rcm), m_waypoint_count(0), m_waypoint_count_inv(0.0) { }
    void setWaypointCount(int waypoint_count);
    int waypointCount() const { return m_waypoint_count; }
    void setEndpoints(const std::vector<double>& start, const std::vector<double>& finish, const std::vector<int>& variables);
    void interpolate(int n, std::vector<double>& state, const std::vector<int>& variables) const;
private:
    const RobotCollisionModel* m_rcm;
    std::vector<double> m_start, m_diffs;
    int m_waypoint_count;
    double m_waypoint_count_inv;
};

void FillMotionInterpolation(const RobotMotionCollisionModel& rmcm,
                             const std::vector<double>& start, const std::vector<double>& finish,
                             const std::vector<int>& variables, double res, MotionInterpolation& motion);

/// smpl/include/smpl/angles.h:45-99
double normalize_angle(double angle);
double shortest_angle_diff(double af, double ai);
double shortest_angle_dist(double af, double ai);

} // namespace oracle

#endif
