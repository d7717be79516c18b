// ORACLE (test infrastructure only -- never linked into the product path).
#include "kdl_model.h"

#include <cmath>
#include <cstdlib>

namespace oracle {

KdlFrame KdlFrame::Identity()
{
    KdlFrame f;
    for (int i = 0; i < 9; ++i) f.M[i] = (i % 4 == 0) ? 1.0 : 0.0;
    f.p[0] = f.p[1] = f.p[2] = 0.0;
    return f;
}

/// KDL frames.inl: Frame(lhs.M*rhs.M, lhs.M*rhs.p + lhs.p)
KdlFrame operator*(const KdlFrame& a, const KdlFrame& b)
{
    KdlFrame r;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) {
            r.M[3 * i + j] = a.M[3 * i + 0] * b.M[0 + j] + a.M[3 * i + 1] * b.M[3 + j] + a.M[3 * i + 2] * b.M[6 + j];
        }
        r.p[i] = (a.M[3 * i + 0] * b.p[0] + a.M[3 * i + 1] * b.p[1] + a.M[3 * i + 2] * b.p[2]) + a.p[i];
    }
    return r;
}

/// KDL Frame::Inverse(): Frame(M^T, -(M^T * p))
static KdlFrame Inverse(const KdlFrame& f)
{
    KdlFrame r;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) {
            r.M[3 * i + j] = f.M[3 * j + i];
        }
    }
    for (int i = 0; i < 3; ++i) {
        r.p[i] = -(r.M[3 * i + 0] * f.p[0] + r.M[3 * i + 1] * f.p[1] + r.M[3 * i + 2] * f.p[2]);
    }
    return r;
}

/// KDL Rotation::Rot2(rotvec, angle) (frames.cpp)
static void Rot2(const double v[3], double angle, double M[9])
{
    double ct = std::cos(angle);
    double st = std::sin(angle);
    double vt = 1 - ct;
    double m_vt_0 = vt * v[0];
    double m_vt_1 = vt * v[1];
    double m_vt_2 = vt * v[2];
    double m_st_0 = v[0] * st;
    double m_st_1 = v[1] * st;
    double m_st_2 = v[2] * st;
    double m_vt_0_1 = m_vt_0 * v[1];
    double m_vt_0_2 = m_vt_0 * v[2];
    double m_vt_1_2 = m_vt_1 * v[2];
    M[0] = ct + m_vt_0 * v[0];   M[1] = -m_st_2 + m_vt_0_1;   M[2] = m_st_1 + m_vt_0_2;
    M[3] = m_st_2 + m_vt_0_1;    M[4] = ct + m_vt_1 * v[1];   M[5] = -m_st_0 + m_vt_1_2;
    M[6] = -m_st_1 + m_vt_0_2;   M[7] = m_st_0 + m_vt_1_2;    M[8] = ct + m_vt_2 * v[2];
}

/// urdf rpy -> quaternion (urdf::Rotation::setFromRPY) -> KDL Rotation::Quaternion
static KdlFrame FrameFromXyzRpy(const double xyz[3], const double rpy[3])
{
    double phi = rpy[0] / 2.0, the = rpy[1] / 2.0, psi = rpy[2] / 2.0;
    double x = std::sin(phi) * std::cos(the) * std::cos(psi) - std::cos(phi) * std::sin(the) * std::sin(psi);
    double y = std::cos(phi) * std::sin(the) * std::cos(psi) + std::sin(phi) * std::cos(the) * std::sin(psi);
    double z = std::cos(phi) * std::cos(the) * std::sin(psi) - std::sin(phi) * std::sin(the) * std::cos(psi);
    double w = std::cos(phi) * std::cos(the) * std::cos(psi) + std::sin(phi) * std::sin(the) * std::sin(psi);
    double s = std::sqrt(x * x + y * y + z * z + w * w);
    if (std::fabs(s) < 1e-5) {
        x = 0.0; y = 0.0; z = 0.0; w = 1.0;
    } else {
        x /= s; y /= s; z /= s; w /= s;
    }
    double x2 = x * x, y2 = y * y, z2 = z * z, w2 = w * w;
    KdlFrame f;
    f.M[0] = w2 + x2 - y2 - z2;       f.M[1] = 2 * x * y - 2 * w * z;   f.M[2] = 2 * x * z + 2 * w * y;
    f.M[3] = 2 * x * y + 2 * w * z;   f.M[4] = w2 - x2 + y2 - z2;       f.M[5] = 2 * y * z - 2 * w * x;
    f.M[6] = 2 * x * z - 2 * w * y;   f.M[7] = 2 * y * z + 2 * w * x;   f.M[8] = w2 - x2 - y2 + z2;
    f.p[0] = xyz[0]; f.p[1] = xyz[1]; f.p[2] = xyz[2];
    return f;
}

/// joint.pose(q) (KDL joint.cpp)
static KdlFrame JointPose(const KDLRobotModel::Segment& s, double q)
{
    KdlFrame f = KdlFrame::Identity();
    if (s.joint_kind == 1) {
        Rot2(s.axis, q, f.M);
        f.p[0] = s.origin[0]; f.p[1] = s.origin[1]; f.p[2] = s.origin[2];
    } else if (s.joint_kind == 2) {
        f.p[0] = s.origin[0] + s.axis[0] * q;
        f.p[1] = s.origin[1] + s.axis[1] * q;
        f.p[2] = s.origin[2] + s.axis[2] * q;
    }
    return f;
}

/// kdl_robot_model.cpp:59-158 (FK + limits only; IK solvers out of scope)
bool KDLRobotModel::init(
    const RobotDesc& desc,
    const std::vector<std::string>& planning_joints,
    const std::string& chain_root_link,
    const std::string& chain_tip_link,
    std::string* err)
{
    m_planning_joints = planning_joints;
    m_T_kin_to_planning = KdlFrame::Identity();
    planning_segment_nr = 0;

    std::map<std::string, const JointDesc*> parent_joint;
    for (const JointDesc& j : desc.joints) {
        parent_joint[j.child] = &j;
    }
    // walk tip -> root (KDL::Tree::getChain)
    std::vector<const JointDesc*> chain;
    std::string link = chain_tip_link;
    while (link != chain_root_link) {
        auto it = parent_joint.find(link);
        if (it == parent_joint.end()) {
            if (err) *err = "Failed to fetch the KDL chain (root: " + chain_root_link + ", tip: " + chain_tip_link + ")";
            return false;
        }
        chain.insert(chain.begin(), it->second);
        link = it->second->parent;
    }

    segments.clear();
    for (const JointDesc* j : chain) {
        Segment s;
        s.name = j->child;
        KdlFrame F = FrameFromXyzRpy(j->xyz, j->rpy);
        s.q_index = -1;
        if (j->type == "fixed") {
            s.joint_kind = 0;
            s.axis[0] = s.axis[1] = s.axis[2] = 0.0;
            s.origin[0] = s.origin[1] = s.origin[2] = 0.0;
        } else {
            s.joint_kind = (j->type == "prismatic") ? 2 : 1;
            // F.M * axis, then Joint ctor normalises
            double a[3];
            for (int i = 0; i < 3; ++i) {
                a[i] = F.M[3 * i + 0] * j->axis[0] + F.M[3 * i + 1] * j->axis[1] + F.M[3 * i + 2] * j->axis[2];
            }
            double n = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
            for (int i = 0; i < 3; ++i) {
                s.axis[i] = a[i] / n;
                s.origin[i] = F.p[i];
            }
        }
        // f_tip = joint.pose(0).Inverse() * F_parent_jnt
        s.f_tip = Inverse(JointPose(s, 0.0)) * F;
        segments.push_back(s);
    }

    // every non-fixed joint of a KDL chain consumes one q entry, in chain
    // order; the planning joints are required to be exactly those joints
    // (kdl_robot_model.cpp:98-110 checks membership; jnt_pos_in_ is sized by
    // getNrOfJoints and filled positionally, :401-404)
    int qi = 0;
    for (Segment& s : segments) {
        if (s.joint_kind != 0) {
            s.q_index = qi++;
        }
    }
    if (qi != (int)planning_joints.size()) {
        if (err) *err = "planning joints do not match the movable joints of the chain";
        return false;
    }
    {
        int k = 0;
        for (const JointDesc* j : chain) {
            if (j->type != "fixed") {
                if (j->name != planning_joints[k]) {
                    if (err) *err = "planning joint order differs from chain order at '" + j->name + "'";
                    return false;
                }
                ++k;
            }
        }
    }

    // joint limits: kdl_robot_model.cpp:236-320
    min_limits.resize(planning_joints.size());
    max_limits.resize(planning_joints.size());
    continuous.resize(planning_joints.size());
    for (size_t i = 0; i < planning_joints.size(); ++i) {
        const JointDesc* jd = nullptr;
        for (const JointDesc* j : chain) {
            if (j->name == planning_joints[i]) jd = j;
        }
        if (!jd) {
            if (err) *err = "Joint limits were not found for " + planning_joints[i];
            return false;
        }
        if (jd->type != "continuous") {
            continuous[i] = false;
            if (!jd->has_safety) {
                min_limits[i] = jd->lower;
                max_limits[i] = jd->upper;
            } else {
                min_limits[i] = jd->soft_lower;
                max_limits[i] = jd->soft_upper;
            }
        } else {
            min_limits[i] = -M_PI;
            max_limits[i] = M_PI;
            continuous[i] = true;
        }
    }
    return true;
}

bool KDLRobotModel::setPlanningLink(const std::string& name)
{
    for (size_t i = 0; i < segments.size(); ++i) {
        if (segments[i].name == name) {
            planning_segment_nr = (int)i; // link_map_[name] = segment index (:152-154)
            return true;
        }
    }
    return false;
}

/// kdl_robot_model.cpp:173-189
double KDLRobotModel::normalizeAngle(double a, double a_min, double a_max) const
{
    if (std::fabs(a) > 2.0 * M_PI) {
        a = std::fmod(a, 2.0 * M_PI);
    }
    while (a > a_max) {
        a -= 2.0 * M_PI;
    }
    while (a < a_min) {
        a += 2.0 * M_PI;
    }
    return a;
}

/// kdl_robot_model.cpp:326-337 -> normalizeAnglesIntoRange :210-235
bool KDLRobotModel::checkJointLimits(const std::vector<double>& angles_in) const
{
    std::vector<double> angles = angles_in;
    size_t dim = angles.size();
    if (min_limits.size() != dim || max_limits.size() != dim) {
        return false;
    }
    for (size_t i = 0; i < dim; i++) {
        if (min_limits[i] > max_limits[i]) {
            return false;
        }
    }
    for (size_t i = 0; i < dim; i++) {
        double min_angle_norm = normalize_angle(min_limits[i]);
        angles[i] = normalizeAngle(angles[i], min_limits[i], min_angle_norm);
        if (angles[i] < min_limits[i] || angles[i] > max_limits[i]) {
            return false;
        }
    }
    return true;
}

/// KDL Rotation::GetRPY (frames.cpp)
static void GetRPY(const double M[9], double& roll, double& pitch, double& yaw)
{
    double epsilon = 1E-12;
    pitch = std::atan2(-M[6], std::sqrt(M[0] * M[0] + M[3] * M[3]));
    if (std::fabs(pitch) > (M_PI / 2.0 - epsilon)) {
        yaw = std::atan2(-M[1], M[4]);
        roll = 0.0;
    } else {
        roll = std::atan2(M[7], M[8]);
        yaw = std::atan2(M[3], M[0]);
    }
}

/// kdl_robot_model.cpp:400-423
bool KDLRobotModel::computePlanningLinkFK(const std::vector<double>& angles, std::vector<double>& pose) const
{
    pose.assign(6, 0.0);
    std::vector<double> q(angles);
    for (size_t i = 0; i < continuous.size() && i < q.size(); ++i) { // normalizeAngles :191-198
        if (continuous[i]) {
            q[i] = normalize_angle(q[i]);
        }
    }
    // ChainFkSolverPos_recursive::JntToCart(q, f1, segmentNr)
    KdlFrame f1 = KdlFrame::Identity();
    for (int i = 0; i < planning_segment_nr; ++i) {
        const Segment& s = segments[i];
        double qq = (s.q_index >= 0) ? q[s.q_index] : 0.0;
        f1 = f1 * (JointPose(s, qq) * s.f_tip);
    }
    KdlFrame f = m_T_kin_to_planning * f1;
    pose[0] = f.p[0];
    pose[1] = f.p[1];
    pose[2] = f.p[2];
    GetRPY(f.M, pose[3], pose[4], pose[5]);
    return true;
}

/// manip_lattice.cpp:2297-2312
std::vector<double> GetTargetOffsetPose(const std::vector<double>& tip_pose, const double xyz_offset[3])
{
    Affine3 T = Translation(tip_pose[0], tip_pose[1], tip_pose[2]);
    T = T * AngleAxis(tip_pose[5], Vec3(0.0, 0.0, 1.0));
    T = T * AngleAxis(tip_pose[4], Vec3(0.0, 1.0, 0.0));
    T = T * AngleAxis(tip_pose[3], Vec3(1.0, 0.0, 0.0));
    T = T * Translation(xyz_offset[0], xyz_offset[1], xyz_offset[2]);
    return { T(0, 3), T(1, 3), T(2, 3), tip_pose[3], tip_pose[4], tip_pose[5] };
}

///////////////////////////////////////////////////////////////////////////////
// BfsHeuristic
///////////////////////////////////////////////////////////////////////////////

BfsHeuristic::BfsHeuristic(const EuclidDistanceMap* grid, double inflation_radius, int cost_per_cell) :
    wall_count(0),
    m_grid(grid),
    m_inflation_radius(inflation_radius),
    m_cost_per_cell(cost_per_cell)
{
    syncGridAndBfs();
}

/// bfs_heuristic.cpp:331-353
void BfsHeuristic::syncGridAndBfs()
{
    const int xc = m_grid->numCellsX();
    const int yc = m_grid->numCellsY();
    const int zc = m_grid->numCellsZ();
    m_bfs.reset(new BFS_3D(xc, yc, zc));
    wall_count = 0;
    for (int x = 0; x < xc; ++x) {
    for (int y = 0; y < yc; ++y) {
    for (int z = 0; z < zc; ++z) {
        const double radius = m_inflation_radius;
        if (m_grid->getDistance(x, y, z) <= radius) {
            m_bfs->setWall(x, y, z);
            ++wall_count;
        }
    }
    }
    }
}

/// bfs_heuristic.cpp:83-101
bool BfsHeuristic::updateGoal(double x, double y, double z)
{
    int gx, gy, gz;
    m_grid->worldToGrid(x, y, z, gx, gy, gz);
    bool in = m_bfs->inBounds(gx, gy, gz);
    m_bfs->run(gx, gy, gz);
    return in;
}

/// bfs_heuristic.cpp:355-366
int BfsHeuristic::getBfsCostToGoal(int x, int y, int z) const
{
    if (!m_bfs->inBounds(x, y, z)) {
        return Infinity;
    } else if (m_bfs->getDistance(x, y, z) == BFS_3D::WALL) {
        return Infinity;
    } else {
        return m_cost_per_cell * m_bfs->getDistance(x, y, z);
    }
}

/// bfs_heuristic.cpp:148-163
int BfsHeuristic::getGoalHeuristicAt(double x, double y, double z) const
{
    int dx, dy, dz;
    m_grid->worldToGrid(x, y, z, dx, dy, dz);
    return getBfsCostToGoal(dx, dy, dz);
}

/// bfs_heuristic.cpp:127-138
double BfsHeuristic::getMetricGoalDistance(double x, double y, double z) const
{
    int gx, gy, gz;
    m_grid->worldToGrid(x, y, z, gx, gy, gz);
    if (!m_bfs->inBounds(gx, gy, gz)) {
        return (double)BFS_3D::WALL * m_grid->resolution();
    } else {
        return (double)m_bfs->getDistance(gx, gy, gz) * m_grid->resolution();
    }
}

} // namespace oracle
