// ORACLE (test infrastructure only -- never linked into the product path).
#include "collision_model.h"

#include <algorithm>
#include <cassert>
#include <cmath>
#include <functional>
#include <limits>

namespace oracle {

///////////////////////////////////////////////////////////////////////////////
// angles.h:45-99
///////////////////////////////////////////////////////////////////////////////

double normalize_angle(double angle)
{
    // normalize to [-2*pi, 2*pi] range
    if (std::fabs(angle) > 2.0 * M_PI) {
        angle = std::fmod(angle, 2.0 * M_PI);
    }
    if (angle < -M_PI) {
        angle += 2.0 * M_PI;
    }
    if (angle > M_PI) {
        angle -= 2.0 * M_PI;
    }
    return angle;
}

double shortest_angle_diff(double af, double ai) { return normalize_angle(af - ai); }
double shortest_angle_dist(double af, double ai) { return std::fabs(shortest_angle_diff(af, ai)); }

///////////////////////////////////////////////////////////////////////////////
// transform_functions.h:95-258
///////////////////////////////////////////////////////////////////////////////

Affine3 ComputeJointTransform(JointFn fn, const Affine3& o, const Vec3& axis, const double* jvals)
{
    switch (fn) {
    case FN_FIXED:
        return o;
    case FN_REV_X: { // transform_functions.h:104-137
        Affine3 t;
        double cth = std::cos(jvals[0]);
        double sth = std::sin(jvals[0]);
        for (int r = 0; r < 3; ++r) {
            t(r, 0) = o(r, 0);
            t(r, 1) = cth * o(r, 1) + sth * o(r, 2);
            t(r, 2) = cth * o(r, 2) - sth * o(r, 1);
            t(r, 3) = o(r, 3);
        }
        return t;
    }
    case FN_REV_Y: { // transform_functions.h:139-172
        Affine3 t;
        double cth = std::cos(jvals[0]);
        double sth = std::sin(jvals[0]);
        for (int r = 0; r < 3; ++r) {
            t(r, 0) = cth * o(r, 0) - sth * o(r, 2);
            t(r, 1) = o(r, 1);
            t(r, 2) = sth * o(r, 0) + cth * o(r, 2);
            t(r, 3) = o(r, 3);
        }
        return t;
    }
    case FN_REV_Z: { // transform_functions.h:174-207
        Affine3 t;
        double cth = std::cos(jvals[0]);
        double sth = std::sin(jvals[0]);
        for (int r = 0; r < 3; ++r) {
            t(r, 0) = o(r, 0) * cth + o(r, 1) * sth;
            t(r, 1) = o(r, 1) * cth - o(r, 0) * sth;
            t(r, 2) = o(r, 2);
            t(r, 3) = o(r, 3);
        }
        return t;
    }
    case FN_REV_GENERIC: // :95-102, :209-216
        return o * AngleAxis(jvals[0], axis);
    case FN_PRISMATIC: // :218-226  (translation along local z regardless of axis)
        return o * Translation(0.0, 0.0, jvals[0]);
    case FN_PLANAR: // :240-249
        return (o * Translation(jvals[0], jvals[1], 0.0)) * AngleAxis(jvals[2], Vec3(0.0, 0.0, 1.0));
    case FN_FLOATING:
        // not reachable in any oracle scene (world joint is fixed or planar);
        // see transform_functions.h:228-238
        return o * Translation(jvals[0], jvals[1], jvals[2]);
    }
    return o;
}

///////////////////////////////////////////////////////////////////////////////
// base_collision_models.cpp
///////////////////////////////////////////////////////////////////////////////

/// base_collision_models.cpp:569-592
void ComputeOptimalBoundingSphere(const SphereModel& s1, const SphereModel& s2, Vec3& c, double& r)
{
    const Vec3& p = s1.center;
    const Vec3& q = s2.center;
    Vec3 v = q - p;
    const double dist = norm(v);
    if (s1.radius > dist + s2.radius) { // s1 contains s2
        c = s1.center;
        r = s1.radius;
    } else if (s2.radius > dist + s1.radius) { // s2 contains s1
        c = s2.center;
        r = s2.radius;
    } else {
        Vec3 vn = normalized(v);
        Vec3 a = q + vn * s2.radius;
        Vec3 b = p - vn * s1.radius;
        c = 0.5 * (a + b);
        r = 0.5 * norm(a - b);
    }
}

/// base_collision_models.cpp:594-641
static int ComputeLargestBoundingBoxAxis(
    std::vector<const SphereConfig*>::iterator first,
    std::vector<const SphereConfig*>::iterator last)
{
    if (std::distance(first, last) == 0) {
        return 0;
    }
    double minx = (*first)->x, miny = (*first)->y, minz = (*first)->z;
    double maxx = minx, maxy = miny, maxz = minz;
    for (auto it = first; it != last; ++it) {
        const SphereConfig& s = **it;
        if (s.x < minx) minx = s.x;
        if (s.y < miny) miny = s.y;
        if (s.z < minz) minz = s.z;
        if (s.x > maxx) maxx = s.x;
        if (s.y > maxy) maxy = s.y;
        if (s.z > maxz) maxz = s.z;
    }
    const double spanx = maxx - minx;
    const double spany = maxy - miny;
    const double spanz = maxz - minz;
    if (spanx > spany && spanx > spanz) {
        return 0;
    } else if (spany > spanz) {
        return 1;
    } else {
        return 2;
    }
}

/// base_collision_models.cpp:184-222
void SphereModelTree::buildFrom(const std::vector<SphereConfig>& spheres)
{
    nodes.clear();
    std::vector<const SphereConfig*> sptrs(spheres.size());
    for (size_t i = 0; i < spheres.size(); ++i) {
        sptrs[i] = &spheres[i];
    }
    buildRecursive(sptrs.begin(), sptrs.end());
}

/// base_collision_models.cpp:337-444
int SphereModelTree::buildRecursive(
    std::vector<const SphereConfig*>::iterator msfirst,
    std::vector<const SphereConfig*>::iterator mslast)
{
    if (mslast == msfirst) {
        return -1;
    }

    const auto count = std::distance(msfirst, mslast);
    if (count == 1) {
        nodes.emplace_back();
        SphereModel& cs = nodes.back();
        const SphereConfig& s = **msfirst;
        cs.name = s.name;
        cs.center = Vec3(s.x, s.y, s.z);
        cs.radius = s.radius;
        cs.priority = s.priority;
        cs.left = -1;
        cs.right = -1;
        return (int)nodes.size() - 1;
    }

    // compute the largest axis of the bounding box along which to split
    const int split_axis = ComputeLargestBoundingBoxAxis(msfirst, mslast);

    // compute the average centroid of all spheres at this level
    Vec3 compact_center(0.0, 0.0, 0.0);
    for (auto it = msfirst; it != mslast; ++it) {
        const SphereConfig& s = **it;
        compact_center = compact_center + Vec3(s.x, s.y, s.z);
    }
    compact_center = compact_center / (double)count;

    // radius required to encompass all model spheres from the centroid
    double compact_radius = 0.0;
    for (auto it = msfirst; it != mslast; ++it) {
        Vec3 c((*it)->x, (*it)->y, (*it)->z);
        double radius = norm(c - compact_center) + (*it)->radius;
        if (radius > compact_radius) {
            compact_radius = radius;
        }
    }

    // split the tree along the largest axis by the centroid (libstdc++
    // std::partition, as the reference)
    std::vector<const SphereConfig*>::iterator msmid;
    if (split_axis == 0) {
        const double x = compact_center.x;
        msmid = std::partition(msfirst, mslast, [x](const SphereConfig* s) { return s->x < x; });
    } else if (split_axis == 1) {
        const double y = compact_center.y;
        msmid = std::partition(msfirst, mslast, [y](const SphereConfig* s) { return s->y < y; });
    } else {
        const double z = compact_center.z;
        msmid = std::partition(msfirst, mslast, [z](const SphereConfig* s) { return s->z < z; });
    }
    if (msfirst == msmid || msmid == mslast) {
        msmid = msfirst + (std::distance(msfirst, mslast) >> 1);
    }

    // recurse on both subtrees
    const int left_idx = buildRecursive(msfirst, msmid);
    const int right_idx = buildRecursive(msmid, mslast);

    // optimal sphere containing both child spheres
    Vec3 greedy_center;
    double greedy_radius;
    ComputeOptimalBoundingSphere(nodes[left_idx], nodes[right_idx], greedy_center, greedy_radius);

    nodes.emplace_back();
    SphereModel& sphere = nodes.back();
    if (greedy_radius < compact_radius) {
        sphere.center = greedy_center;
        sphere.radius = greedy_radius;
    } else {
        sphere.center = compact_center;
        sphere.radius = compact_radius;
    }
    sphere.priority = 0;
    sphere.left = left_idx;
    sphere.right = right_idx;
    return (int)nodes.size() - 1;
}

///////////////////////////////////////////////////////////////////////////////
// RobotCollisionModel  (robot_collision_model.cpp)
///////////////////////////////////////////////////////////////////////////////

static bool set_err(std::string* err, const std::string& msg)
{
    if (err) {
        *err = msg;
    }
    return false;
}

/// robot_collision_model.cpp:288-516
void RobotCollisionModel::addJoint(const JointDesc* j, const std::string& jname, JointType type)
{
    joint_names.push_back(jname);
    if (j) {
        joint_origins.push_back(FromXyzRpy(j->xyz[0], j->xyz[1], j->xyz[2], j->rpy[0], j->rpy[1], j->rpy[2]));
        joint_axes.push_back(Vec3(j->axis[0], j->axis[1], j->axis[2]));
    } else {
        joint_origins.push_back(Affine3::Identity());
        joint_axes.push_back(Vec3(0.0, 0.0, 0.0));
    }
    joint_var_indices.emplace_back((int)jvar_names.size(), 0);

    const double inf = std::numeric_limits<double>::infinity();
    auto add_var = [&](const std::string& n, bool continuous, bool bounded, double lo, double hi) {
        jvar_names.push_back(n);
        jvar_continuous.push_back(continuous);
        jvar_has_position_bounds.push_back(bounded);
        jvar_min_positions.push_back(lo);
        jvar_max_positions.push_back(hi);
        jvar_joint_indices.push_back((int)joint_types.size());
        jvar_name_to_index[n] = (int)jvar_names.size() - 1;
    };
    auto rev_fn = [&]() {
        const Vec3& a = joint_axes.back();
        if (a.x == 1.0 && a.y == 0.0 && a.z == 0.0) return FN_REV_X;
        if (a.x == 0.0 && a.y == 1.0 && a.z == 0.0) return FN_REV_Y;
        if (a.x == 0.0 && a.y == 0.0 && a.z == 1.0) return FN_REV_Z;
        return FN_REV_GENERIC;
    };

    switch (type) {
    case FIXED:
        joint_fns.push_back(FN_FIXED);
        joint_types.push_back(FIXED);
        break;
    case REVOLUTE:
    case PRISMATIC: {
        double lo = j->has_safety ? j->soft_lower : j->lower;
        double hi = j->has_safety ? j->soft_upper : j->upper;
        add_var(jname, false, j->has_limits, lo, hi);
        joint_types.push_back(type);
        joint_fns.push_back(type == REVOLUTE ? rev_fn() : FN_PRISMATIC);
    }   break;
    case CONTINUOUS:
        add_var(jname, true, false, -inf, inf);
        joint_types.push_back(CONTINUOUS);
        joint_fns.push_back(rev_fn());
        break;
    case PLANAR:
        add_var(jname + "/x", false, false, -inf, inf);
        add_var(jname + "/y", false, false, -inf, inf);
        add_var(jname + "/theta", true, false, -inf, inf);
        joint_types.push_back(PLANAR);
        joint_fns.push_back(FN_PLANAR);
        break;
    case FLOATING:
        add_var(jname + "/trans_x", false, false, -inf, inf);
        add_var(jname + "/trans_y", false, false, -inf, inf);
        add_var(jname + "/trans_z", true, false, -inf, inf);
        add_var(jname + "/rot_x", false, true, -1.0, 1.0);
        add_var(jname + "/rot_y", false, true, -1.0, 1.0);
        add_var(jname + "/rot_z", true, true, -1.0, 1.0);
        add_var(jname + "/rot_w", true, true, -1.0, 1.0);
        joint_types.push_back(FLOATING);
        joint_fns.push_back(FN_FLOATING);
        break;
    }
    joint_var_indices.back().second = (int)jvar_names.size();
}

static bool ParseJointType(const std::string& s, JointType& t)
{
    if (s == "fixed") t = FIXED;
    else if (s == "revolute") t = REVOLUTE;
    else if (s == "prismatic") t = PRISMATIC;
    else if (s == "continuous") t = CONTINUOUS;
    else if (s == "planar") t = PLANAR;
    else if (s == "floating") t = FLOATING;
    else return false;
    return true;
}

bool RobotCollisionModel::init(const RobotDesc& desc, std::string* err)
{
    // ---- initRobotModel: robot_collision_model.cpp:117-286 ----
    name = desc.name;
    model_frame = desc.root;

    // urdf-style maps: link -> parent joint, link -> child joints in
    // joint-name order (urdfdom initTree walks a std::map keyed by joint name)
    std::map<std::string, const JointDesc*> joints_by_name;
    for (const JointDesc& j : desc.joints) {
        joints_by_name[j.name] = &j;
    }
    std::map<std::string, const JointDesc*> parent_joint;
    std::map<std::string, std::vector<const JointDesc*>> child_joints;
    for (const auto& e : joints_by_name) {
        const JointDesc* j = e.second;
        if (parent_joint.count(j->child)) {
            return set_err(err, "link '" + j->child + "' has two parent joints");
        }
        parent_joint[j->child] = j;
        child_joints[j->parent].push_back(j);
    }

    JointType wtype;
    if (!ParseJointType(desc.world_joint_type, wtype) ||
        !(wtype == FLOATING || wtype == PLANAR || wtype == FIXED))
    {
        return set_err(err, "World joint config has invalid type");
    }
    addJoint(nullptr, desc.world_joint_name, wtype);
    joint_parent_links.push_back(-1);

    // depth-first traversal using an explicit stack, children pushed in
    // reverse so they are visited in child_joints order
    std::vector<std::string> stack;
    stack.push_back(desc.root);
    while (!stack.empty()) {
        std::string link = stack.back();
        stack.pop_back();

        link_names.push_back(link);
        const int lidx = (int)link_names.size() - 1;
        link_name_to_index[link] = lidx;
        link_children_joints.push_back(std::vector<int>());

        auto pit = parent_joint.find(link);
        if (pit != parent_joint.end()) {
            const JointDesc* pj = pit->second;
            JointType t;
            if (!ParseJointType(pj->type, t)) {
                return set_err(err, "Unknown joint type '" + pj->type + "'");
            }
            addJoint(pj, pj->name, t);
            joint_parent_links.push_back(link_name_to_index[pj->parent]);
            link_parent_joints.push_back((int)joint_names.size() - 1);
        } else {
            link_parent_joints.push_back(0);
        }

        auto cit = child_joints.find(link);
        if (cit != child_joints.end()) {
            for (auto it = cit->second.rbegin(); it != cit->second.rend(); ++it) {
                stack.push_back((*it)->child);
            }
        }
    }

    // map joint -> child link; link -> child joints
    joint_child_links.resize(joint_names.size());
    for (size_t lidx = 0; lidx < link_names.size(); ++lidx) {
        joint_child_links[link_parent_joints[lidx]] = (int)lidx;
    }
    for (size_t jidx = 0; jidx < joint_names.size(); ++jidx) {
        int plidx = joint_parent_links[jidx];
        if (plidx >= 0) {
            link_children_joints[plidx].push_back((int)jidx);
        }
    }

    // ---- initCollisionModel: robot_collision_model.cpp:517-623 ----
    std::vector<GroupConfig> expanded_groups;
    if (!expandGroups(desc.groups, expanded_groups, err)) {
        return false;
    }

    spheres_models.reserve(desc.spheres_models.size());
    for (const SpheresModelConfig& cfg : desc.spheres_models) {
        if (cfg.spheres.empty()) {
            continue;
        }
        if (!hasLink(cfg.link_name)) {
            return set_err(err, "spheres model for unknown link '" + cfg.link_name + "'");
        }
        spheres_models.push_back(SpheresModel());
        spheres_models.back().spheres.buildFrom(cfg.spheres);
        spheres_models.back().link_index = linkIndex(cfg.link_name);
    }

    // voxels models.  The reference voxelises the URDF collision geometry of
    // the link (robot_collision_model.cpp:959-1076); the oracle's fixture
    // gives one axis-aligned box per link instead (OUR stand-in for meshes):
    // voxel centres on a res-spaced lattice filling the box.
    voxels_models.resize(desc.voxels_models.size());
    for (size_t i = 0; i < voxels_models.size(); ++i) {
        const VoxelsModelConfig& cfg = desc.voxels_models[i];
        if (!hasLink(cfg.link_name)) {
            return set_err(err, "voxels model for unknown link '" + cfg.link_name + "'");
        }
        VoxelsModel& vm = voxels_models[i];
        vm.link_index = linkIndex(cfg.link_name);
        vm.voxel_res = cfg.res;
        int n[3];
        for (int a = 0; a < 3; ++a) {
            n[a] = (int)std::floor(cfg.size[a] / cfg.res + 0.5);
        }
        if (n[0] > 0 && n[1] > 0 && n[2] > 0) {
            for (int ix = 0; ix < n[0]; ++ix) {
            for (int iy = 0; iy < n[1]; ++iy) {
            for (int iz = 0; iz < n[2]; ++iz) {
                vm.voxels.push_back(Vec3(
                    cfg.center[0] - 0.5 * cfg.size[0] + (ix + 0.5) * cfg.res,
                    cfg.center[1] - 0.5 * cfg.size[1] + (iy + 0.5) * cfg.res,
                    cfg.center[2] - 0.5 * cfg.size[2] + (iz + 0.5) * cfg.res));
            }
            }
            }
        }
    }

    group_models.resize(expanded_groups.size());
    for (size_t i = 0; i < group_models.size(); ++i) {
        group_models[i].name = expanded_groups[i].name;
        for (const std::string& ln : expanded_groups[i].links) {
            if (!hasLink(ln)) {
                return set_err(err, "group '" + group_models[i].name + "' references unknown link '" + ln + "'");
            }
            group_models[i].link_indices.push_back(linkIndex(ln));
        }
        group_name_to_index[group_models[i].name] = (int)i;
    }

    link_spheres_models.assign(link_names.size(), -1);
    for (size_t i = 0; i < spheres_models.size(); ++i) {
        link_spheres_models[spheres_models[i].link_index] = (int)i;
    }
    link_voxels_models.assign(link_names.size(), -1);
    for (size_t i = 0; i < voxels_models.size(); ++i) {
        link_voxels_models[voxels_models[i].link_index] = (int)i;
    }
    return true;
}

/// robot_collision_model.cpp:625-756
bool RobotCollisionModel::expandGroups(
    const std::vector<GroupConfig>& groups,
    std::vector<GroupConfig>& expanded_groups,
    std::string* err) const
{
    std::vector<GroupConfig> expanded;
    for (const GroupConfig& g : groups) {
        GroupConfig config;
        config.name = g.name;
        config.links = g.links;
        expanded.push_back(config);
    }

    // expand chains
    for (size_t gidx = 0; gidx < groups.size(); ++gidx) {
        for (const auto& chain : groups[gidx].chains) {
            const std::string& base = chain.first;
            const std::string& tip = chain.second;
            std::vector<std::string> chain_links;
            std::string link_name = tip;
            chain_links.push_back(link_name);
            while (link_name != base) {
                if (!hasLink(link_name)) {
                    return set_err(err, "link '" + link_name + "' not found in the robot model");
                }
                int lidx = linkIndex(link_name);
                int pjidx = link_parent_joints[lidx];
                int plidx = joint_parent_links[pjidx];
                if (plidx < 0) {
                    return set_err(err, "(" + base + ", " + tip + ") is not a chain in the robot model");
                }
                link_name = link_names[plidx];
                chain_links.push_back(link_name);
            }
            expanded[gidx].links.insert(expanded[gidx].links.end(), chain_links.begin(), chain_links.end());
        }
    }

    // expand subgroups in dependency order
    auto name_to_index = [&](const std::string& n) {
        for (size_t i = 0; i < groups.size(); ++i) {
            if (groups[i].name == n) return i;
        }
        return groups.size();
    };
    std::vector<std::vector<size_t>> deps(groups.size()), rdeps(groups.size());
    std::vector<size_t> waiting(groups.size(), 0);
    for (size_t gidx = 0; gidx < groups.size(); ++gidx) {
        for (const std::string& dep : groups[gidx].groups) {
            size_t didx = name_to_index(dep);
            if (didx == groups.size()) {
                return set_err(err, "group '" + groups[gidx].name + "' depends on unknown group '" + dep + "'");
            }
            deps[gidx].push_back(didx);
            rdeps[didx].push_back(gidx);
        }
        waiting[gidx] = groups[gidx].groups.size();
    }
    std::vector<size_t> q(groups.size());
    for (size_t i = 0; i < q.size(); ++i) q[i] = i;
    while (!q.empty()) {
        auto git = std::find_if(q.begin(), q.end(), [&](size_t i) { return waiting[i] == 0; });
        if (git == q.end()) {
            return set_err(err, "cycle in group config");
        }
        size_t gidx = *git;
        --waiting[gidx];
        q.erase(git);
        for (size_t didx : deps[gidx]) {
            expanded[gidx].links.insert(
                expanded[gidx].links.end(), expanded[didx].links.begin(), expanded[didx].links.end());
        }
        for (size_t rdidx : rdeps[gidx]) {
            --waiting[rdidx];
        }
    }

    for (GroupConfig& eg : expanded) {
        std::sort(eg.links.begin(), eg.links.end());
        eg.links.erase(std::unique(eg.links.begin(), eg.links.end()), eg.links.end());
    }
    expanded_groups = expanded;
    return true;
}

///////////////////////////////////////////////////////////////////////////////
// RobotCollisionState  (robot_collision_state.{h,cpp})
///////////////////////////////////////////////////////////////////////////////

/// robot_collision_state.cpp:207-239 (initRobotState), :241-330 (initCollisionState)
RobotCollisionState::RobotCollisionState(const RobotCollisionModel* model) :
    link_transform_updates(0),
    m_model(model)
{
    m_jvar_positions.assign(model->jointVarCount(), 0.0);
    for (int vidx = 0; vidx < model->jointVarCount(); ++vidx) {
        if (model->jvar_has_position_bounds[vidx]) {
            if (model->jvar_min_positions[vidx] > 0.0 || model->jvar_max_positions[vidx] < 0.0) {
                m_jvar_positions[vidx] = 0.5 * (model->jvar_min_positions[vidx] + model->jvar_max_positions[vidx]);
            }
        }
    }
    m_dirty_joint_transforms.assign(model->jointCount(), true);
    m_joint_transforms.assign(model->jointCount(), Affine3::Identity());
    m_dirty_link_transforms.assign(model->linkCount(), true);
    m_link_transforms.assign(model->linkCount(), Affine3::Identity());
    m_link_transform_versions.assign(model->linkCount(), -1);
    m_dirty_link_transforms[0] = false;
    m_link_transform_versions[0] = 0;

    m_sphere_states.resize(model->spheres_models.size());
    for (size_t i = 0; i < model->spheres_models.size(); ++i) {
        const SphereModelTree& tree = model->spheres_models[i].spheres;
        m_sphere_states[i].resize(tree.nodes.size());
        for (size_t s = 0; s < tree.nodes.size(); ++s) {
            m_sphere_states[i][s].pos = tree.nodes[s].center; // base_collision_states.cpp:58
        }
    }

    m_dirty_voxels_states.assign(model->voxels_models.size(), true);
    m_voxels_states.resize(model->voxels_models.size());
    for (size_t i = 0; i < model->voxels_models.size(); ++i) {
        m_voxels_states[i] = model->voxels_models[i].voxels; // link-frame copies until first update
    }

    m_group_spheres_indices.resize(model->group_models.size());
    m_group_voxels_indices.resize(model->group_models.size());
    for (size_t ssidx = 0; ssidx < model->spheres_models.size(); ++ssidx) {
        const int lidx = model->spheres_models[ssidx].link_index;
        for (size_t gidx = 0; gidx < model->group_models.size(); ++gidx) {
            const auto& li = model->group_models[gidx].link_indices;
            if (std::find(li.begin(), li.end(), lidx) != li.end()) {
                m_group_spheres_indices[gidx].push_back((int)ssidx);
            }
        }
    }
    for (size_t vsidx = 0; vsidx < model->voxels_models.size(); ++vsidx) {
        const int lidx = model->voxels_models[vsidx].link_index;
        for (size_t gidx = 0; gidx < model->group_models.size(); ++gidx) {
            const auto& li = model->group_models[gidx].link_indices;
            if (std::find(li.begin(), li.end(), lidx) == li.end()) {
                m_group_voxels_indices[gidx].push_back((int)vsidx);
            }
        }
    }
    m_link_voxels_states.assign(model->linkCount(), -1);
    for (size_t i = 0; i < model->voxels_models.size(); ++i) {
        m_link_voxels_states[model->voxels_models[i].link_index] = (int)i;
    }
}

/// robot_collision_state.h:211-302.  Only the cases the oracle scenes use:
/// FIXED (identity) and PLANAR / FLOATING with a pure-translation / identity
/// rotation argument.
bool RobotCollisionState::setWorldToModelTransform(const Affine3& transform)
{
    bool updated = false;
    Affine3 M = Affine3::Identity();
    switch (m_model->joint_types[0]) {
    case FIXED:
        break;
    case PLANAR: {
        double x = transform(0, 3);
        double y = transform(1, 3);
        double theta = 0.0; // identity rotation: s_squared < 10 eps -> 0
        updated |= (m_jvar_positions[0] != x);
        updated |= (m_jvar_positions[1] != y);
        updated |= (m_jvar_positions[2] != theta);
        if (updated) {
            m_jvar_positions[0] = x;
            m_jvar_positions[1] = y;
            m_jvar_positions[2] = theta;
            M = Translation(x, y, 0.0) * AngleAxis(theta, Vec3(0.0, 0.0, 1.0));
        }
    }   break;
    default:
        assert(!"unsupported world joint type in oracle");
        break;
    }
    if (updated) {
        m_link_transforms[0] = M;
        std::fill(m_dirty_link_transforms.begin(), m_dirty_link_transforms.end(), true);
        m_dirty_link_transforms[0] = false;
        ++m_link_transform_versions[0];
        std::fill(m_dirty_voxels_states.begin(), m_dirty_voxels_states.end(), true);
        return true;
    }
    return false;
}

/// robot_collision_state.cpp:58-103
bool RobotCollisionState::setJointVarPosition(int vidx, double position)
{
    if (m_jvar_positions[vidx] != position) {
        m_jvar_positions[vidx] = position;
        const int jidx = m_model->jvar_joint_indices[vidx];
        m_dirty_joint_transforms[jidx] = true;
        std::vector<int>& q = m_q;
        q.clear();
        q.push_back(m_model->joint_child_links[jidx]);
        while (!q.empty()) {
            int lidx = q.back();
            q.pop_back();
            m_dirty_link_transforms[lidx] = true;
            if (m_link_voxels_states[lidx] >= 0) {
                m_dirty_voxels_states[m_link_voxels_states[lidx]] = true;
            }
            for (int cjidx : m_model->link_children_joints[lidx]) {
                q.push_back(m_model->joint_child_links[cjidx]);
            }
        }
        return true;
    }
    return false;
}

/// robot_collision_state.cpp:105-166.  The "keep only the most ancestral
/// joints" filtering is an optimisation of the dirtying walk; dirtying from
/// every changed joint marks exactly the same set of links.
bool RobotCollisionState::setJointVarPositions(const double* positions)
{
    bool any = false;
    for (size_t vidx = 0; vidx < m_jvar_positions.size(); ++vidx) {
        any |= setJointVarPosition((int)vidx, positions[vidx]);
    }
    return any;
}

/// robot_collision_state.h:385-431
bool RobotCollisionState::updateLinkTransform(int lidx)
{
    if (!m_dirty_link_transforms[lidx]) {
        return false;
    }
    assert(lidx != 0);
    int pjidx = m_model->link_parent_joints[lidx];
    int plidx = m_model->joint_parent_links[pjidx];
    if (plidx >= 0) {
        updateLinkTransform(plidx);
    }
    if (m_dirty_joint_transforms[pjidx]) {
        int fvidx = m_model->joint_var_indices[pjidx].first;
        const double* variables = m_jvar_positions.data() + fvidx;
        m_joint_transforms[pjidx] = ComputeJointTransform(
            m_model->joint_fns[pjidx], m_model->joint_origins[pjidx], m_model->joint_axes[pjidx], variables);
        m_dirty_joint_transforms[pjidx] = false;
    }
    const Affine3& T_parent_link = m_joint_transforms[pjidx];
    if (plidx >= 0) {
        m_link_transforms[lidx] = m_link_transforms[plidx] * T_parent_link;
    } else {
        m_link_transforms[lidx] = T_parent_link;
    }
    m_dirty_link_transforms[lidx] = false;
    ++m_link_transform_versions[lidx];
    ++link_transform_updates;
    return true;
}

/// robot_collision_state.h:560-581
bool RobotCollisionState::updateSphereState(int ssidx, int sidx)
{
    const int lidx = m_model->spheres_models[ssidx].link_index;
    const int link_version = m_link_transform_versions[lidx];
    SphereState& sphere_state = m_sphere_states[ssidx][sidx];
    if (!m_dirty_link_transforms[lidx] && sphere_state.version == link_version) {
        return false;
    }
    updateLinkTransform(lidx);
    sphere_state.pos = m_link_transforms[lidx] * m_model->spheres_models[ssidx].spheres.nodes[sidx].center;
    sphere_state.version = m_link_transform_versions[lidx];
    return true;
}

/// robot_collision_state.h:472-497
bool RobotCollisionState::updateVoxelsState(int vsidx)
{
    if (!m_dirty_voxels_states[vsidx]) {
        return false;
    }
    const VoxelsModel& vm = m_model->voxels_models[vsidx];
    const int lidx = vm.link_index;
    if (lidx != 0) {
        updateLinkTransform(lidx);
    }
    const Affine3& T_model_link = m_link_transforms[lidx];
    std::vector<Vec3> new_voxels(vm.voxels.size());
    for (size_t i = 0; i < vm.voxels.size(); ++i) {
        new_voxels[i] = T_model_link * vm.voxels[i];
    }
    m_voxels_states[vsidx] = new_voxels;
    m_dirty_voxels_states[vsidx] = false;
    return true;
}

///////////////////////////////////////////////////////////////////////////////
// RobotMotionCollisionModel  (robot_motion_collision_model.cpp:41-275)
///////////////////////////////////////////////////////////////////////////////

RobotMotionCollisionModel::RobotMotionCollisionModel(const RobotCollisionModel* rcm) : m_rcm(rcm)
{
    const int nj = rcm->jointCount();
    std::vector<int> q_joint(nj, -1);
    int q_head = 0, q_tail = 0;
    std::vector<int> p_joint(nj, 0);

    for (int lidx = 0; lidx < rcm->linkCount(); ++lidx) {
        if (rcm->link_children_joints[lidx].empty()) {
            q_joint[q_tail++] = rcm->link_parent_joints[lidx];
        }
    }

    std::vector<Vec3> mc(nj);
    std::vector<double> mr(nj, 0.0);
    std::vector<Vec3> mrc(nj);
    std::vector<double> mrr(nj);
    std::vector<std::vector<Vec3>> sample_spheres(nj);
    std::vector<double> sample_radii(nj, 0.0);

    while (q_head != q_tail) {
        int jidx = q_joint[q_head++];

        std::vector<Vec3> joint_frame_sample_centers;
        std::vector<double> joint_frame_sample_radii;

        // R(n)
        const int clidx = rcm->joint_child_links[jidx];
        if (rcm->hasSpheresModel(clidx)) {
            const SphereModelTree& tree = rcm->spheres_models[rcm->link_spheres_models[clidx]].spheres;
            const SphereModel& root = tree.nodes[tree.root()];
            joint_frame_sample_centers.push_back(root.center);
            joint_frame_sample_radii.push_back(root.radius);
        }

        // M_samples(n+1)
        for (int cjidx : rcm->link_children_joints[clidx]) {
            if (sample_radii[cjidx] != 0.0) {
                const Affine3& T_joint_child = rcm->joint_origins[cjidx];
                for (const Vec3& sphere_pos : sample_spheres[cjidx]) {
                    joint_frame_sample_centers.push_back(T_joint_child * sphere_pos);
                    joint_frame_sample_radii.push_back(sample_radii[cjidx]);
                }
            }
        }

        // MR(n)
        Vec3 mr_center(0.0, 0.0, 0.0);
        double mr_radius = 0.0;
        if (!joint_frame_sample_centers.empty()) {
            for (size_t i = 0; i < joint_frame_sample_centers.size(); ++i) {
                mr_center = mr_center + joint_frame_sample_centers[i];
            }
            mr_center = mr_center / (double)joint_frame_sample_centers.size();
            for (size_t i = 0; i < joint_frame_sample_centers.size(); ++i) {
                double radius = norm(joint_frame_sample_centers[i] - mr_center) + joint_frame_sample_radii[i];
                mr_radius = std::max(mr_radius, radius);
            }
        }
        mrc[jidx] = mr_center;
        mrr[jidx] = mr_radius;

        // M(n): sample the joint variable, update MR, compute M
        std::vector<Vec3> samples;
        if (mr_radius != 0.0) {
            double res = 2.0 * M_PI / 180.0;
            const JointType jt = rcm->joint_types[jidx];
            if (jt == REVOLUTE) {
                int vidx = rcm->joint_var_indices[jidx].first;
                double span = rcm->jvar_max_positions[vidx] - rcm->jvar_min_positions[vidx];
                int sample_count = (int)std::round(span / res) + 1;
                res = span / (sample_count - 1);
                for (int i = 0; i < sample_count; ++i) {
                    double alpha = (double)i / (double)(sample_count - 1);
                    double val = (1.0 - alpha) * rcm->jvar_min_positions[vidx] + alpha * rcm->jvar_max_positions[vidx];
                    const Affine3 T_joint_link = AngleAxis(val, rcm->joint_axes[jidx]);
                    samples.push_back(T_joint_link * mr_center);
                }
            } else if (jt == CONTINUOUS) {
                int sample_count = (int)std::round(2.0 * M_PI / res);
                res = 2.0 * M_PI / sample_count;
                for (int i = 0; i < sample_count; ++i) {
                    double val = i * res;
                    const Affine3 T_joint_link = AngleAxis(val, rcm->joint_axes[jidx]);
                    samples.push_back(T_joint_link * mr_center);
                }
            } else if (jt == PRISMATIC) {
                int vidx = rcm->joint_var_indices[jidx].first;
                double span = rcm->jvar_max_positions[vidx] - rcm->jvar_min_positions[vidx];
                int sample_count = (int)std::round(span / res) + 1;
                res = span / (sample_count - 1);
                for (int i = 0; i < sample_count; ++i) {
                    double alpha = (double)i / (double)(sample_count - 1);
                    double val = (1.0 - alpha) * rcm->jvar_min_positions[vidx] + alpha * rcm->jvar_max_positions[vidx];
                    const Vec3 t = val * rcm->joint_axes[jidx];
                    const Affine3 T_joint_link = Translation(t.x, t.y, t.z);
                    samples.push_back(T_joint_link * mr_center);
                }
            } else if (jt == FIXED) {
                samples.push_back(rcm->joint_origins[jidx] * mr_center);
            }
            // FLOATING / PLANAR: "TODO: Cannot sample" in the reference -> no samples
        }

        sample_spheres[jidx] = samples;
        sample_radii[jidx] = mr_radius;

        // M(n)
        Vec3 m_center(0.0, 0.0, 0.0);
        double m_radius = 0.0;
        if (!samples.empty()) {
            for (const Vec3& c : samples) {
                m_center = m_center + c;
            }
            m_center = m_center / (double)samples.size();
            for (const Vec3& c : samples) {
                m_radius = std::max(m_radius, norm(c - m_center) + mr_radius);
            }
        }
        mc[jidx] = m_center;
        mr[jidx] = m_radius;

        int plidx = rcm->joint_parent_links[jidx];
        if (plidx >= 0) {
            int pjidx = rcm->link_parent_joints[plidx];
            if (pjidx >= 0) {
                ++p_joint[pjidx];
                if (p_joint[pjidx] == (int)rcm->link_children_joints[plidx].size()) {
                    q_joint[q_tail++] = pjidx;
                }
            }
        }
    }

    m_centers = mc;
    m_radii = mr;
    mr_centers = mrc;
    mr_radii = mrr;
}

/// robot_motion_collision_model.cpp:371-407
double RobotMotionCollisionModel::getMaxSphereMotion(
    const std::vector<double>& start,
    const std::vector<double>& finish,
    const std::vector<int>& variables) const
{
    double motion = 0.0;
    for (size_t i = 0; i < start.size(); ++i) {
        const int vidx = variables[i];
        const int jidx = m_rcm->jvar_joint_indices[vidx];
        double dist = 0.0;
        switch (m_rcm->joint_types[jidx]) {
        case FIXED:
            break;
        case CONTINUOUS:
            dist = shortest_angle_dist(finish[i], start[i]);
            motion += (norm(mr_centers[jidx]) + mr_radii[jidx]) * dist;
            break;
        case REVOLUTE:
            dist = std::fabs(finish[i] - start[i]);
            motion += (norm(mr_centers[jidx]) + mr_radii[jidx]) * dist;
            break;
        case PRISMATIC:
            dist = std::fabs(finish[i] - start[i]);
            motion += dist;
            break;
        case PLANAR:
        case FLOATING:
            break;
        }
    }
    return motion;
}

/// robot_motion_collision_model.h:173-181
void MotionInterpolation::setWaypointCount(int waypoint_count)
{
    if (waypoint_count) {
        m_waypoint_count = std::max(2, waypoint_count);
        m_waypoint_count_inv = 1.0 / (double)(m_waypoint_count - 1);
    } else {
        m_waypoint_count = waypoint_count;
    }
}

/// robot_motion_collision_model.h:224-249
void MotionInterpolation::setEndpoints(
    const std::vector<double>& start,
    const std::vector<double>& finish,
    const std::vector<int>& variables)
{
    m_start = start;
    m_diffs.resize(variables.size());
    for (size_t vidx = 0; vidx < variables.size(); ++vidx) {
        int jidx = m_rcm->jvar_joint_indices[variables[vidx]];
        switch (m_rcm->joint_types[jidx]) {
        case FIXED:
            break;
        case REVOLUTE:
        case PRISMATIC:
            m_diffs[vidx] = finish[vidx] - start[vidx];
            break;
        case CONTINUOUS:
            m_diffs[vidx] = shortest_angle_diff(finish[vidx], start[vidx]);
            break;
        case PLANAR:
        case FLOATING:
            break;
        }
    }
}

/// robot_motion_collision_model.h:297-321
void MotionInterpolation::interpolate(int n, std::vector<double>& state, const std::vector<int>& variables) const
{
    state.resize(m_start.size());
    const double alpha = (double)n * m_waypoint_count_inv;
    for (size_t v = 0; v < variables.size(); ++v) {
        int jidx = m_rcm->jvar_joint_indices[variables[v]];
        switch (m_rcm->joint_types[jidx]) {
        case FIXED:
            break;
        case REVOLUTE:
        case CONTINUOUS:
        case PRISMATIC:
            state[v] = m_start[v] + alpha * m_diffs[v];
            break;
        case PLANAR:
        case FLOATING:
            break;
        }
    }
}

/// robot_motion_collision_model.h:352-366
void FillMotionInterpolation(
    const RobotMotionCollisionModel& rmcm,
    const std::vector<double>& start,
    const std::vector<double>& finish,
    const std::vector<int>& variables,
    double res,
    MotionInterpolation& motion)
{
    motion.setEndpoints(start, finish, variables);
    double max_motion = rmcm.getMaxSphereMotion(start, finish, variables);
    if (max_motion == 0.0) {
        motion.setWaypointCount(0);
    } else {
        motion.setWaypointCount((int)std::ceil(max_motion / res) + 1);
    }
}

} // namespace oracle
