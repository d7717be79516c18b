// ORACLE (test infrastructure only -- never linked into the product path).
//
// Flat C API over the oracle classes so tests/ (ctypes), __graft_entry__.smoke()
// and bench.py's cpu_baseline leg can drive it.  Nothing under smpl_b200/ may
// load this library.
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "bfs3d.h"
#include "collision_space.h"
#include "distance_map.h"
#include "kdl_model.h"
#include "arastar.h"
#include "lattice.h"
#include "shortcut.h"
#include "voxelize.h"
#include "robot_desc.h"

using namespace oracle;

static thread_local std::string g_err;

struct oracle_scene
{
    RobotDesc desc;
    std::string group;
    std::vector<std::string> planning_joints;
    std::unique_ptr<EuclidDistanceMap> df;
    std::unique_ptr<OccupancyGrid> grid;
    std::unique_ptr<CollisionSpace> cc;
    std::unique_ptr<KDLRobotModel> kdl;
    std::unique_ptr<BfsHeuristic> heur;
    double xyz_offset[3];
    int dof;
    oracle_scene() : dof(0) { xyz_offset[0] = xyz_offset[1] = xyz_offset[2] = 0.0; }
};

static std::vector<std::string> split_csv(const char* s)
{
    std::vector<std::string> out;
    std::stringstream ss(s ? s : "");
    std::string item;
    while (std::getline(ss, item, ',')) {
        if (!item.empty()) out.push_back(item);
    }
    return out;
}

extern "C" {

const char* oracle_last_error(void) { return g_err.c_str(); }

oracle_scene* oracle_scene_create(
    const char* robot_path, const char* group, const char* planning_joints_csv,
    const double* origin, const double* size, double res, double max_dist)
{
    std::unique_ptr<oracle_scene> s(new oracle_scene);
    if (!LoadRobotDesc(robot_path, s->desc, &g_err)) {
        return nullptr;
    }
    s->group = group;
    s->planning_joints = split_csv(planning_joints_csv);
    s->dof = (int)s->planning_joints.size();
    s->df.reset(new EuclidDistanceMap(origin[0], origin[1], origin[2], size[0], size[1], size[2], res, max_dist));
    s->grid.reset(new OccupancyGrid(s->df.get()));
    s->cc.reset(new CollisionSpace);
    if (!s->cc->init(s->grid.get(), s->desc, s->group, s->planning_joints, &g_err)) {
        return nullptr;
    }
    return s.release();
}

void oracle_scene_destroy(oracle_scene* s) { delete s; }

int oracle_scene_dof(oracle_scene* s) { return s->dof; }

int oracle_scene_set_joint(oracle_scene* s, const char* name, double value)
{
    return s->cc->setJointPosition(name, value) ? 0 : -1;
}

/// call_planner.cpp:441-1527 -> CollisionSpace::setAllowedCollisionMatrix
int oracle_scene_use_desc_acm(oracle_scene* s)
{
    AllowedCollisionMatrix acm;
    for (const AcmEntryDesc& e : s->desc.acm) {
        acm.setEntry(e.a, e.b, e.allowed);
    }
    s->cc->setAllowedCollisionMatrix(acm);
    return 0;
}

int oracle_scene_acm_set(oracle_scene* s, const char* a, const char* b, int allowed)
{
    AllowedCollisionMatrix acm = s->cc->acm();
    acm.setEntry(a, b, allowed != 0);
    s->cc->setAllowedCollisionMatrix(acm);
    return 0;
}

int oracle_scene_set_padding(oracle_scene* s, double padding)
{
    s->cc->setPadding(padding);
    return 0;
}

/// occupied cells given as effective grid coordinates (x,y,z triples)
int oracle_scene_add_cells(oracle_scene* s, const int32_t* xyz, int n)
{
    std::vector<std::array<int, 3>> cells(n);
    for (int i = 0; i < n; ++i) {
        cells[i] = {{ xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2] }};
    }
    s->df->addCellsToMap(cells);
    return 0;
}

int oracle_scene_add_points(oracle_scene* s, const double* xyz, int n)
{
    std::vector<Vec3> pts(n);
    for (int i = 0; i < n; ++i) {
        pts[i] = Vec3(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    }
    s->df->addPointsToMap(pts);
    return 0;
}

int oracle_scene_remove_points(oracle_scene* s, const double* xyz, int n)
{
    std::vector<Vec3> pts(n);
    for (int i = 0; i < n; ++i) {
        pts[i] = Vec3(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    }
    s->df->removePointsFromMap(pts);
    return 0;
}

int oracle_scene_attach_spheres(oracle_scene* s, const char* id, const char* link, const double* centers, int n, double radius)
{
    std::vector<Vec3> c(n);
    for (int i = 0; i < n; ++i) {
        c[i] = Vec3(centers[3 * i], centers[3 * i + 1], centers[3 * i + 2]);
    }
    return s->cc->attachSpheres(id, c, radius, link) ? 0 : -1;
}

/// AttachedBodiesCollisionModel::attachBody for a box shape: generateSpheresModel
/// (attached_bodies_collision_model.cpp:264-309) -- VoxelizeShape at 0.025 / sqrt(2), voxel origin zero, one
/// sphere of radius 0.025 per surface voxel.  Returns the number of spheres.
int oracle_scene_attach_box(oracle_scene* s, const char* id, const char* link, const double* size, const double* pose3x4)
{
    const double object_enclosing_sphere_radius = 0.025;
    const double zero[3] = { 0.0, 0.0, 0.0 };
    Affine3 pose;
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 4; ++c) pose.m[r][c] = pose3x4[4 * r + c];
    }
    std::vector<Vec3> voxels;
    VoxelizeBox(size[0], size[1], size[2], pose, object_enclosing_sphere_radius / std::sqrt(2), zero, false, voxels);
    if (!s->cc->attachSpheres(id, voxels, object_enclosing_sphere_radius, link)) {
        return -1;
    }
    return (int)voxels.size();
}

int oracle_scene_detach(oracle_scene* s, const char* id)
{
    return s->cc->detachBody(id) ? 0 : -1;
}

/// the first collision check inserts the voxels of out-of-group links into the
/// distance field (self_collision_model.cpp:616-837); afterwards the field is
/// static for a fixed set of non-planning joint values
int oracle_scene_prime(oracle_scene* s, const double* q)
{
    std::vector<double> st(q, q + s->dof);
    (void)s->cc->isStateValid(st);
    return 0;
}

void oracle_scene_grid_info(oracle_scene* s, int32_t* dims, double* origin, double* res, int32_t* dmax_sq)
{
    dims[0] = s->df->numCellsX();
    dims[1] = s->df->numCellsY();
    dims[2] = s->df->numCellsZ();
    origin[0] = s->df->originX();
    origin[1] = s->df->originY();
    origin[2] = s->df->originZ();
    *res = s->df->resolution();
    *dmax_sq = s->df->dmaxSqrd();
}

/// integer squared cell distances, unpadded, x-major / z-fastest
int oracle_scene_df_d2(oracle_scene* s, int32_t* out)
{
    const int nx = s->df->numCellsX(), ny = s->df->numCellsY(), nz = s->df->numCellsZ();
    size_t k = 0;
    for (int x = 0; x < nx; ++x) {
    for (int y = 0; y < ny; ++y) {
    for (int z = 0; z < nz; ++z) {
        out[k++] = s->df->getSquaredCellDistance(x, y, z);
    }
    }
    }
    return 0;
}

int oracle_world_to_grid(oracle_scene* s, const double* xyz, int n, int32_t* out)
{
    for (int i = 0; i < n; ++i) {
        int x, y, z;
        s->df->worldToGrid(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], x, y, z);
        out[3 * i] = x; out[3 * i + 1] = y; out[3 * i + 2] = z;
    }
    return 0;
}

///////////////////////////////////////////////////////////////////////////////
// validity
///////////////////////////////////////////////////////////////////////////////

int oracle_is_states_valid(oracle_scene* s, const double* q, int n, uint8_t* verdict)
{
    std::vector<double> st(s->dof);
    for (int i = 0; i < n; ++i) {
        st.assign(q + (size_t)i * s->dof, q + (size_t)(i + 1) * s->dof);
        verdict[i] = s->cc->isStateValid(st) ? 1 : 0;
    }
    return 0;
}

int oracle_collision_distance(oracle_scene* s, const double* q, int n, double* out)
{
    std::vector<double> st(s->dof);
    for (int i = 0; i < n; ++i) {
        st.assign(q + (size_t)i * s->dof, q + (size_t)(i + 1) * s->dof);
        out[i] = s->cc->collisionDistance(st);
    }
    return 0;
}

int oracle_report_states(oracle_scene* s, const double* q, int n, uint8_t* verdict,
                         int32_t* required_lookups, double* cell_margin, double* pair_margin)
{
    std::vector<double> st(s->dof);
    for (int i = 0; i < n; ++i) {
        st.assign(q + (size_t)i * s->dof, q + (size_t)(i + 1) * s->dof);
        StateReport r = s->cc->reportState(st);
        verdict[i] = r.valid ? 1 : 0;
        if (required_lookups) required_lookups[i] = r.required_lookups;
        if (cell_margin) cell_margin[i] = r.min_cell_boundary_margin;
        if (pair_margin) pair_margin[i] = r.min_sphere_pair_margin;
    }
    return 0;
}

int oracle_is_edges_valid(oracle_scene* s, const double* q0, const double* q1, int n, uint8_t* verdict, int32_t* counts)
{
    std::vector<double> a(s->dof), b(s->dof);
    for (int i = 0; i < n; ++i) {
        a.assign(q0 + (size_t)i * s->dof, q0 + (size_t)(i + 1) * s->dof);
        b.assign(q1 + (size_t)i * s->dof, q1 + (size_t)(i + 1) * s->dof);
        int c = 0;
        verdict[i] = s->cc->isStateToStateValid(a, b, &c) ? 1 : 0;
        if (counts) counts[i] = c;
    }
    return 0;
}

int oracle_report_edges(oracle_scene* s, const double* q0, const double* q1, int n, uint8_t* verdict,
                        int32_t* counts, int32_t* required_lookups)
{
    std::vector<double> a(s->dof), b(s->dof);
    for (int i = 0; i < n; ++i) {
        a.assign(q0 + (size_t)i * s->dof, q0 + (size_t)(i + 1) * s->dof);
        b.assign(q1 + (size_t)i * s->dof, q1 + (size_t)(i + 1) * s->dof);
        int c = 0, L = 0;
        verdict[i] = s->cc->isStateToStateValidExhaustive(a, b, &c, &L) ? 1 : 0;
        if (counts) counts[i] = c;
        if (required_lookups) required_lookups[i] = L;
    }
    return 0;
}

/// waypoints of one edge: out holds max_wp*dof doubles; returns the count
int oracle_edge_waypoints(oracle_scene* s, const double* q0, const double* q1, double* out, int max_wp)
{
    std::vector<double> a(q0, q0 + s->dof), b(q1, q1 + s->dof);
    std::vector<std::vector<double>> wps;
    s->cc->edgeWaypoints(a, b, wps);
    int n = (int)wps.size();
    for (int i = 0; i < n && i < max_wp; ++i) {
        std::copy(wps[i].begin(), wps[i].end(), out + (size_t)i * s->dof);
    }
    return n;
}

int oracle_num_nodes(oracle_scene* s)
{
    std::vector<Vec3> c;
    std::vector<double> st(s->dof, 0.0);
    // sphereCenters size is state independent
    const auto& cc = *s->cc;
    int n = 0;
    for (int ssidx : s->cc->state().groupSpheresStateIndices(cc.groupIndex())) {
        n += (int)cc.model().spheres_models[ssidx].spheres.nodes.size();
    }
    for (int b : cc.groupAttachedBodies()) {
        n += (int)cc.attachedBodies()[b].spheres.nodes.size();
    }
    return n;
}

int oracle_sphere_centers(oracle_scene* s, const double* q, int n, double* out)
{
    std::vector<double> st(s->dof);
    std::vector<Vec3> c;
    size_t k = 0;
    for (int i = 0; i < n; ++i) {
        st.assign(q + (size_t)i * s->dof, q + (size_t)(i + 1) * s->dof);
        s->cc->sphereCenters(st, c);
        for (const Vec3& v : c) {
            out[k++] = v.x; out[k++] = v.y; out[k++] = v.z;
        }
    }
    return 0;
}

/// model tables in group order, for comparison with the product's builder.
/// per node: cx cy cz radius left right tree_index link_index (8 doubles)
int oracle_node_table(oracle_scene* s, double* out)
{
    const auto& cc = *s->cc;
    size_t k = 0;
    int tree = 0;
    for (int ssidx : s->cc->state().groupSpheresStateIndices(cc.groupIndex())) {
        const auto& sm = cc.model().spheres_models[ssidx];
        for (const SphereModel& n : sm.spheres.nodes) {
            out[k++] = n.center.x; out[k++] = n.center.y; out[k++] = n.center.z; out[k++] = n.radius;
            out[k++] = n.left; out[k++] = n.right; out[k++] = tree; out[k++] = sm.link_index;
        }
        ++tree;
    }
    for (int b : cc.groupAttachedBodies()) {
        const auto& ab = cc.attachedBodies()[b];
        for (const SphereModel& n : ab.spheres.nodes) {
            out[k++] = n.center.x; out[k++] = n.center.y; out[k++] = n.center.z; out[k++] = n.radius;
            out[k++] = n.left; out[k++] = n.right; out[k++] = tree; out[k++] = ab.link_index;
        }
        ++tree;
    }
    return (int)(k / 8);
}

/// per planning variable: ||MR_center|| + MR_radius of its joint (the weight
/// getMaxSphereMotion multiplies |dq| by), joint type
int oracle_motion_weights(oracle_scene* s, double* weights, int32_t* types)
{
    const auto& cc = *s->cc;
    for (int i = 0; i < s->dof; ++i) {
        int vidx = cc.planningVariables()[i];
        int jidx = cc.model().jvar_joint_indices[vidx];
        weights[i] = norm(cc.motionModel().mr_centers[jidx]) + cc.motionModel().mr_radii[jidx];
        types[i] = (int)cc.model().joint_types[jidx];
    }
    return 0;
}

/// number of checked robot tree pairs (after priming), pairs as group-tree indices
int oracle_checked_pairs(oracle_scene* s, int32_t* out, int max_pairs)
{
    const auto& cc = *s->cc;
    const auto& gss = s->cc->state().groupSpheresStateIndices(cc.groupIndex());
    auto tree_of = [&](int ssidx) {
        return (int)(std::find(gss.begin(), gss.end(), ssidx) - gss.begin());
    };
    int n = 0;
    for (const auto& p : cc.checkedSpheresStates()) {
        if (n < max_pairs) {
            out[2 * n] = tree_of(p.first);
            out[2 * n + 1] = tree_of(p.second);
        }
        ++n;
    }
    for (const auto& p : cc.checkedAttachedAttached()) {
        if (n < max_pairs) {
            out[2 * n] = (int)gss.size() + p.first;
            out[2 * n + 1] = (int)gss.size() + p.second;
        }
        ++n;
    }
    for (const auto& p : cc.checkedAttachedRobot()) {
        if (n < max_pairs) {
            out[2 * n] = (int)gss.size() + p.first;
            out[2 * n + 1] = tree_of(p.second);
        }
        ++n;
    }
    return n;
}

void oracle_scene_stats(oracle_scene* s, int64_t* df_lookups, int64_t* pair_tests, int64_t* link_updates)
{
    *df_lookups = s->cc->stats.df_lookups;
    *pair_tests = s->cc->stats.sphere_pair_tests;
    *link_updates = s->cc->state().link_transform_updates;
}

///////////////////////////////////////////////////////////////////////////////
// KDL-like planning model + BFS heuristic
///////////////////////////////////////////////////////////////////////////////

int oracle_scene_init_kdl(oracle_scene* s, const char* chain_root, const char* chain_tip,
                          const char* planning_link, const double* T_kin_to_planning_3x4,
                          const double* xyz_offset)
{
    s->kdl.reset(new KDLRobotModel);
    if (!s->kdl->init(s->desc, s->planning_joints, chain_root, chain_tip, &g_err)) {
        s->kdl.reset();
        return -1;
    }
    if (!s->kdl->setPlanningLink(planning_link)) {
        g_err = "planning link not in chain";
        return -1;
    }
    KdlFrame f;
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) f.M[3 * r + c] = T_kin_to_planning_3x4[4 * r + c];
        f.p[r] = T_kin_to_planning_3x4[4 * r + 3];
    }
    s->kdl->setKinematicsToPlanningTransform(f);
    for (int i = 0; i < 3; ++i) s->xyz_offset[i] = xyz_offset ? xyz_offset[i] : 0.0;
    return 0;
}

int oracle_check_joint_limits(oracle_scene* s, const double* q, int n, uint8_t* out)
{
    std::vector<double> st(s->dof);
    for (int i = 0; i < n; ++i) {
        st.assign(q + (size_t)i * s->dof, q + (size_t)(i + 1) * s->dof);
        out[i] = s->kdl->checkJointLimits(st) ? 1 : 0;
    }
    return 0;
}

void oracle_joint_limits(oracle_scene* s, double* mins, double* maxs, uint8_t* continuous)
{
    for (int i = 0; i < s->dof; ++i) {
        mins[i] = s->kdl->min_limits[i];
        maxs[i] = s->kdl->max_limits[i];
        continuous[i] = s->kdl->continuous[i] ? 1 : 0;
    }
}

/// computePlanningFrameFK: planning link FK then target offset; pose6 per state
int oracle_planning_frame_fk(oracle_scene* s, const double* q, int n, double* pose6)
{
    std::vector<double> st(s->dof), pose;
    for (int i = 0; i < n; ++i) {
        st.assign(q + (size_t)i * s->dof, q + (size_t)(i + 1) * s->dof);
        s->kdl->computePlanningLinkFK(st, pose);
        pose = GetTargetOffsetPose(pose, s->xyz_offset);
        std::copy(pose.begin(), pose.end(), pose6 + (size_t)i * 6);
    }
    return 0;
}

int oracle_heur_init(oracle_scene* s, double inflation_radius, int cost_per_cell)
{
    s->heur.reset(new BfsHeuristic(s->df.get(), inflation_radius, cost_per_cell));
    return s->heur->wall_count;
}

int oracle_heur_set_goal(oracle_scene* s, double x, double y, double z)
{
    return s->heur->updateGoal(x, y, z) ? 0 : 1;
}

int oracle_heur_grid(oracle_scene* s, int32_t* out)
{
    const auto& g = s->heur->bfs()->grid();
    std::copy(g.begin(), g.end(), out);
    return (int)g.size();
}

int oracle_goal_heuristics(oracle_scene* s, const double* q, int n, int32_t* h)
{
    std::vector<double> st(s->dof), pose;
    for (int i = 0; i < n; ++i) {
        st.assign(q + (size_t)i * s->dof, q + (size_t)(i + 1) * s->dof);
        s->kdl->computePlanningLinkFK(st, pose);
        pose = GetTargetOffsetPose(pose, s->xyz_offset);
        h[i] = s->heur->getGoalHeuristicAt(pose[0], pose[1], pose[2]);
    }
    return 0;
}

///////////////////////////////////////////////////////////////////////////////
// planning query: ManipLattice + ARA* over the oracle's checker / heuristic
///////////////////////////////////////////////////////////////////////////////

// per-primitive action weights for the next oracle_plan calls on this thread (empty = 1 each)
static thread_local std::vector<double> tl_prim_weights;

// set by oracle_plan_lazy around a call of oracle_plan
static thread_local bool tl_plan_lazy = false;
static thread_local int tl_plan_lazy_evaluations = 0;

/// mprims: n_prims rows of dof deltas (radians); short_flags[n_prims].
/// out_summary: success, expansions, cost, path_len, num_states.  Returns seconds spent in plan().
double oracle_plan(oracle_scene* s, const double* start, const double* goal_xyz,
                   const double* resolutions, const double* mprims, const uint8_t* short_flags, int n_prims,
                   int use_short_dist, double short_dist_thresh, double epsilon, int max_expansions,
                   const double* xyz_tolerance, int32_t* out_summary, int32_t* path_ids, int max_path,
                   double* path_states /* nullable: [max_path][dof], ManipLattice::extractPath */)
{
    PlanParams pp;
    pp.resolutions.assign(resolutions, resolutions + s->dof);
    for (int p = 0; p < n_prims; ++p) {
        MotionPrim mp;
        mp.delta.assign(mprims + (size_t)p * s->dof, mprims + (size_t)(p + 1) * s->dof);
        mp.short_dist = short_flags[p] != 0;
        mp.weight = (int)tl_prim_weights.size() == n_prims ? tl_prim_weights[p] : 1.0;
        pp.mprims.push_back(mp);
    }
    pp.use_short_dist = use_short_dist != 0;
    pp.short_dist_thresh = short_dist_thresh;
    pp.epsilon = epsilon;
    pp.max_expansions = max_expansions;
    for (int i = 0; i < 3; ++i) pp.xyz_tolerance[i] = xyz_tolerance[i];
    ManipLatticePlanner planner(s->cc.get(), s->kdl.get(), s->heur.get(), s->xyz_offset, 0, pp);
    std::vector<double> st(start, start + s->dof);
    auto t0 = std::chrono::steady_clock::now();
    PlanResult r = tl_plan_lazy ? planner.planLazy(st, goal_xyz) : planner.plan(st, goal_xyz);
    tl_plan_lazy_evaluations = r.evaluations;
    auto t1 = std::chrono::steady_clock::now();
    out_summary[0] = r.success ? 1 : 0;
    out_summary[1] = r.expansions;
    out_summary[2] = r.cost;
    out_summary[3] = (int)r.path_ids.size();
    out_summary[4] = r.num_states;
    for (size_t i = 0; i < r.path_ids.size() && (int)i < max_path; ++i) {
        path_ids[i] = r.path_ids[i];
    }
    if (path_states) {
        for (size_t i = 0; i < r.path_states.size() && (int)i < max_path; ++i) {
            std::copy(r.path_states[i].begin(), r.path_states[i].end(), path_states + i * (size_t)s->dof);
        }
        out_summary[5] = (int)r.path_states.size();
    }
    return std::chrono::duration<double>(t1 - t0).count();
}

/// action weights of the primitives (n = n_prims of the following oracle_plan calls; n = 0 clears them)
void oracle_set_prim_weights(const double* weights, int n)
{
    tl_prim_weights.assign(weights, weights + (n > 0 ? n : 0));
}

/// oracle_plan through the lazy successors and oracle/lazy_arastar.h (ManipLatticePlanner::planLazy);
/// oracle_last_lazy_evaluations() = GetTrueCost calls of the last query on this thread
double oracle_plan_lazy(oracle_scene* s, const double* start, const double* goal_xyz,
                        const double* resolutions, const double* mprims, const uint8_t* short_flags, int n_prims,
                        int use_short_dist, double short_dist_thresh, double epsilon, int max_expansions,
                        const double* xyz_tolerance, int32_t* out_summary, int32_t* path_ids, int max_path,
                        double* path_states)
{
    tl_plan_lazy = true;
    const double secs = oracle_plan(s, start, goal_xyz, resolutions, mprims, short_flags, n_prims, use_short_dist,
                                    short_dist_thresh, epsilon, max_expansions, xyz_tolerance, out_summary, path_ids,
                                    max_path, path_states);
    tl_plan_lazy = false;
    return secs;
}

int oracle_last_lazy_evaluations(void)
{
    return tl_plan_lazy_evaluations;
}

/// oracle/arastar.h on an explicit graph (CSR successor lists), same signature as ref_arastar_search
/// (oracle/ref_arastar_shim.cpp): out[0] = found, out[1] = cost, out[2] = expansions, out[3] = path length
int oracle_arastar_search(int n, const int* off, const int* dst, const int* cost, const int* h,
                          int start, int goal, double eps, int max_expansions,
                          int* path, int max_path, int* out)
{
    (void)n;
    AraStar search(
        [&](int s, std::vector<int>& succs, std::vector<int>& costs) {
            for (int k = off[s]; k < off[s + 1]; ++k) {
                succs.push_back(dst[k]);
                costs.push_back(cost[k]);
            }
        },
        [&](int s) { return h[s]; });
    const AraStar::Result r = search.search(start, goal, eps, max_expansions);
    out[0] = r.found ? 1 : 0;
    out[1] = r.cost;
    out[2] = r.expansions;
    out[3] = (int)r.path.size();
    for (int i = 0; i < (int)r.path.size() && i < max_path; ++i) {
        path[i] = r.path[i];
    }
    return 0;
}

///////////////////////////////////////////////////////////////////////////////
// scene ingest: oracle/voxelize.h
///////////////////////////////////////////////////////////////////////////////

static int emit_voxels(const std::vector<Vec3>& voxels, double* out, int max_out)
{
    if ((int)voxels.size() > max_out) {
        return -(int)voxels.size();
    }
    for (size_t i = 0; i < voxels.size(); ++i) {
        out[3 * i] = voxels[i].x;
        out[3 * i + 1] = voxels[i].y;
        out[3 * i + 2] = voxels[i].z;
    }
    return (int)voxels.size();
}

static Affine3 to_affine(const double* m /*3x4 row-major*/)
{
    Affine3 t;
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 4; ++c) t.m[r][c] = m[4 * r + c];
    }
    return t;
}

/// same signatures as oracle/ref_voxelize_shim.cpp
int oracle_voxelize_mesh(const double* vertices, int nv, const int32_t* triangles, int nt, double res,
                         const double* voxel_origin, int fill, double* out, int max_out)
{
    std::vector<Vec3> v(nv);
    for (int i = 0; i < nv; ++i) v[i] = Vec3(vertices[3 * i], vertices[3 * i + 1], vertices[3 * i + 2]);
    std::vector<int> t(triangles, triangles + 3 * (size_t)nt);
    std::vector<Vec3> voxels;
    VoxelizeMesh(v, t, res, voxel_origin, fill != 0, voxels);
    return emit_voxels(voxels, out, max_out);
}

int oracle_voxelize_box(double length, double width, double height, const double* pose3x4, double res,
                        const double* voxel_origin, int fill, double* out, int max_out)
{
    std::vector<Vec3> voxels;
    VoxelizeBox(length, width, height, to_affine(pose3x4), res, voxel_origin, fill != 0, voxels);
    return emit_voxels(voxels, out, max_out);
}

void oracle_box_mesh(double length, double width, double height, double* vertices, int32_t* indices)
{
    std::vector<Vec3> v;
    std::vector<int> t;
    CreateIndexedBoxMesh(length, width, height, v, t);
    for (size_t i = 0; i < v.size(); ++i) {
        vertices[3 * i] = v[i].x;
        vertices[3 * i + 1] = v[i].y;
        vertices[3 * i + 2] = v[i].z;
    }
    std::copy(t.begin(), t.end(), indices);
}

/// CreateIndexed{Box,Sphere,Cylinder,Cone}Mesh at the origin (kind 0..3; dims = l,w,h / r / r,length / r,height);
/// same signature as ref_shape_mesh.  Returns the vertex count, *n_indices the index count.
int oracle_shape_mesh(int kind, const double* dims, double* vertices, int32_t* indices, int* n_indices)
{
    std::vector<Vec3> v;
    std::vector<int> t;
    if (kind == 0) CreateIndexedBoxMesh(dims[0], dims[1], dims[2], v, t);
    else if (kind == 1) CreateIndexedSphereMesh(dims[0], 7, 8, v, t);
    else if (kind == 2) CreateIndexedCylinderMesh(dims[0], dims[1], v, t);
    else CreateIndexedConeMesh(dims[0], dims[1], v, t);
    for (size_t i = 0; i < v.size(); ++i) {
        vertices[3 * i] = v[i].x;
        vertices[3 * i + 1] = v[i].y;
        vertices[3 * i + 2] = v[i].z;
    }
    std::copy(t.begin(), t.end(), indices);
    *n_indices = (int)t.size();
    return (int)v.size();
}

/// WorldCollisionModel::insertObject (world_collision_model.cpp:193-234) for box primitives: VoxelizeBox with the
/// grid origin as voxel origin, fill = false (voxel_operations.cpp:357-369), then addPointsToField.
/// boxes[n][3 + 12] = size, pose 3x4.  Returns the number of voxels handed to the field.
int oracle_scene_insert_boxes(oracle_scene* s, const double* boxes, int n)
{
    int32_t dims[3];
    double origin[3], res;
    int32_t dmax_sq;
    oracle_scene_grid_info(s, dims, origin, &res, &dmax_sq);
    int total = 0;
    for (int i = 0; i < n; ++i) {
        const double* b = boxes + 15 * (size_t)i;
        std::vector<Vec3> voxels;
        VoxelizeBox(b[0], b[1], b[2], to_affine(b + 3), res, origin, false, voxels);
        s->df->addPointsToMap(voxels);
        total += (int)voxels.size();
    }
    return total;
}

///////////////////////////////////////////////////////////////////////////////
// path post-processing: oracle/shortcut.h
///////////////////////////////////////////////////////////////////////////////

/// The two shortcut templates on an index path with explicit tables (valid[n][n], pair_cost[n][n]); same
/// signature as ref_shortcut_path's result (oracle/ref_shortcut_shim.cpp).  algo 0 = greedy, 1 = divide & conquer.
int oracle_shortcut_table(int n, const double* costs, const uint8_t* valid, const double* pair_cost, int algo,
                          int granularity, int32_t* out, int max_out)
{
    std::vector<double> cvec(costs, costs + (n > 0 ? n - 1 : 0));
    std::vector<ShortcutGenerator> gens = {
        [&](int a, int b, double& c) {
            if (!valid[(size_t)a * n + b]) return false;
            c = pair_cost[(size_t)a * n + b];
            return true;
        } };
    std::vector<int> result;
    const bool ok = algo == 0 ? ShortcutPath(n, cvec, gens, result, (size_t)granularity)
                              : DivideAndConquerShortcutPath(n, cvec, gens, result);
    if (!ok || (int)result.size() > max_out) {
        return -1;
    }
    std::copy(result.begin(), result.end(), out);
    return (int)result.size();
}

/// ShortcutPath(rm, cc, pin, pout, type) (post_processing.cpp:284-365) over the oracle's CollisionSpace, one
/// isStateToStateValid per generator request like the reference.  continuous[dof] = !RobotModel::hasPosLimit.
/// Returns the number of output points; out_idx = their indices in the input path; edge_checks (nullable) =
/// number of isStateToStateValid calls made.
int oracle_shortcut_path(oracle_scene* s, const double* path, int n, const uint8_t* continuous, int kind,
                         int32_t* out_idx, int64_t* edge_checks)
{
    std::vector<bool> cont(continuous, continuous + s->dof);
    int64_t checks = 0;
    auto valid = [&](int a, int b) {
        ++checks;
        std::vector<double> q0(path + (size_t)a * s->dof, path + (size_t)(a + 1) * s->dof);
        std::vector<double> q1(path + (size_t)b * s->dof, path + (size_t)(b + 1) * s->dof);
        return s->cc->isStateToStateValid(q0, q1, nullptr);
    };
    const std::vector<int> idx = ShortcutJointPath(cont, path, n, valid, kind);
    std::copy(idx.begin(), idx.end(), out_idx);
    if (edge_checks) *edge_checks = checks;
    return (int)idx.size();
}

/// InterpolatePath (post_processing.cpp:476-540): every segment is replaced by the waypoints of
/// CollisionChecker::interpolatePath when all of them are valid, else its end point is kept.  The reference's
/// interpolatePath rejects every motion whose end points are WITHIN the limits (collision_space.cpp:592-597,
/// SURVEY.md section 8a defect 7); that test is left out here and in the product.
/// out: [max_points][dof]; returns the number of points or -1 when max_points is too small.
int oracle_interpolate_path(oracle_scene* s, const double* path, int n, double* out, int max_points)
{
    const int dof = s->dof;
    std::vector<std::vector<double>> opath;
    if (n > 0) {
        opath.emplace_back(path, path + dof);
    }
    for (int i = 0; i + 1 < n; ++i) {
        std::vector<double> q0(path + (size_t)i * dof, path + (size_t)(i + 1) * dof);
        std::vector<double> q1(path + (size_t)(i + 1) * dof, path + (size_t)(i + 2) * dof);
        std::vector<std::vector<double>> ipath;
        s->cc->edgeWaypoints(q0, q1, ipath);
        bool collision = false;
        for (const std::vector<double>& p : ipath) {
            if (!s->cc->isStateValid(p)) {
                collision = true;
                break;
            }
        }
        if (collision) {
            opath.push_back(q1);
            continue;
        }
        if (!ipath.empty()) {
            opath.insert(opath.end(), ipath.begin() + 1, ipath.end());
        }
    }
    if ((int)opath.size() > max_points) {
        return -1;
    }
    for (size_t i = 0; i < opath.size(); ++i) {
        std::copy(opath[i].begin(), opath[i].end(), out + i * (size_t)dof);
    }
    return (int)opath.size();
}

///////////////////////////////////////////////////////////////////////////////
// stand-alone BFS_3D
///////////////////////////////////////////////////////////////////////////////

BFS_3D* oracle_bfs_create(int nx, int ny, int nz) { return new BFS_3D(nx, ny, nz); }
void oracle_bfs_destroy(BFS_3D* b) { delete b; }

/// walls: one byte per cell, x-fastest unpadded (index = (z*ny + y)*nx + x)
int oracle_bfs_set_walls(BFS_3D* b, const uint8_t* walls)
{
    const int nx = b->dimX() - 2, ny = b->dimY() - 2, nz = b->dimZ() - 2;
    for (int z = 0; z < nz; ++z) {
    for (int y = 0; y < ny; ++y) {
    for (int x = 0; x < nx; ++x) {
        if (walls[((size_t)z * ny + y) * nx + x]) {
            b->setWall(x, y, z);
        }
    }
    }
    }
    return 0;
}

int oracle_bfs_run(BFS_3D* b, int x, int y, int z) { return b->run(x, y, z); }
int oracle_bfs_run_multi(BFS_3D* b, const int32_t* xyz, int count) { return b->run(xyz, count); }

int oracle_bfs_grid(BFS_3D* b, int32_t* out)
{
    std::copy(b->grid().begin(), b->grid().end(), out);
    return (int)b->grid().size();
}

int oracle_bfs_get_distance(BFS_3D* b, int x, int y, int z)
{
    if (!b->inBounds(x, y, z)) {
        return -2; // undefined in the reference
    }
    return b->getDistance(x, y, z);
}

///////////////////////////////////////////////////////////////////////////////
// timing helpers for bench.py's cpu_baseline (time only the checks, as
// benchmark_cc.cpp:243-249)
///////////////////////////////////////////////////////////////////////////////

double oracle_time_states_valid(oracle_scene* s, const double* q, int n, uint8_t* verdict)
{
    auto t0 = std::chrono::steady_clock::now();
    oracle_is_states_valid(s, q, n, verdict);
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

double oracle_time_edges_valid(oracle_scene* s, const double* q0, const double* q1, int n, uint8_t* verdict, int32_t* counts)
{
    auto t0 = std::chrono::steady_clock::now();
    oracle_is_edges_valid(s, q0, q1, n, verdict, counts);
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

double oracle_time_bfs_run(BFS_3D* b, int x, int y, int z)
{
    auto t0 = std::chrono::steady_clock::now();
    b->run(x, y, z);
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

} // extern "C"
