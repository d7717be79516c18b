// Flat C entry points over the host-side C++ (model builder + adapters) so the
// Python plumbing (tests, bench.py) can drive it with ctypes.  Declared in
// include/smplhost.h.
#include <cstring>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/smplgpu.h"
#include "../../include/smplhost.h"
#include "robot_tables.h"

using smplhost::RobotTables;

static thread_local std::string g_err;

struct smplhost_tables
{
    RobotTables t;
    std::vector<double> voxels;
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

const char* smplhost_last_error(void) { return g_err.c_str(); }

smplhost_tables* smplhost_tables_load(const char* robot_path)
{
    std::unique_ptr<smplhost_tables> h(new smplhost_tables);
    if (!robot_path || !h->t.load(robot_path, &g_err)) {
        return nullptr;
    }
    return h.release();
}

void smplhost_tables_destroy(smplhost_tables* h) { delete h; }

int smplhost_tables_configure(smplhost_tables* h, const char* group, const char* planning_joints_csv)
{
    if (!h || !group) return -1;
    return h->t.configure(group, split_csv(planning_joints_csv), &g_err) ? 0 : -1;
}

int smplhost_tables_set_joint(smplhost_tables* h, const char* variable, double value)
{
    if (!h || !variable) return -1;
    if (!h->t.setJointPosition(variable, value)) {
        g_err = std::string("joint variable '") + variable + "' not found";
        return -1;
    }
    return 0;
}

int smplhost_tables_use_file_acm(smplhost_tables* h)
{
    if (!h) return -1;
    h->t.useFileAcm();
    return 0;
}

int smplhost_tables_set_acm_entry(smplhost_tables* h, const char* a, const char* b, int allowed)
{
    if (!h || !a || !b) return -1;
    h->t.setAcmEntry(a, b, allowed != 0);
    return 0;
}

int smplhost_tables_attach_spheres(smplhost_tables* h, const char* id, const char* link,
                                   const double* centers, int n, double radius)
{
    if (!h || !id || !link || !centers) return -1;
    if (!h->t.attachSpheres(id, link, centers, n, radius)) {
        g_err = "attach failed (unknown link, duplicate id or empty body)";
        return -1;
    }
    return 0;
}

int smplhost_tables_detach(smplhost_tables* h, const char* id)
{
    if (!h || !id) return -1;
    return h->t.detach(id) ? 0 : -1;
}

int smplhost_tables_set_planning_chain(smplhost_tables* h, const char* root, const char* tip,
                                       const char* planning_link, const double* T_kin_to_planning,
                                       const double* xyz_offset)
{
    if (!h || !root || !tip || !planning_link) return -1;
    return h->t.setPlanningChain(root, tip, planning_link, T_kin_to_planning, xyz_offset, &g_err) ? 0 : -1;
}

int smplhost_tables_dof(smplhost_tables* h) { return h ? h->t.dof() : -1; }

int smplhost_tables_limits(smplhost_tables* h, double* mins, double* maxs, uint8_t* continuous)
{
    if (!h) return -1;
    for (int i = 0; i < h->t.dof(); ++i) {
        mins[i] = h->t.varMin()[i];
        maxs[i] = h->t.varMax()[i];
        continuous[i] = (uint8_t)h->t.varContinuous()[i];
    }
    return 0;
}

const smplgpu_robot_desc* smplhost_tables_desc(smplhost_tables* h) { return h ? h->t.desc() : nullptr; }

int smplhost_tables_apply(smplhost_tables* h, smplgpu_ctx* ctx)
{
    if (!h || !ctx) return -1;
    int r = smplgpu_set_robot(ctx, h->t.desc());
    if (r != 0) {
        g_err = smplgpu_last_error(ctx);
    }
    return r;
}

int smplhost_tables_outside_voxels(smplhost_tables* h, const double** xyz)
{
    if (!h) return -1;
    h->voxels = h->t.outsideGroupVoxels();
    if (xyz) *xyz = h->voxels.data();
    return (int)(h->voxels.size() / 3);
}

/* per node: cx cy cz radius left right tree link (8 doubles), for comparison with the oracle */
int smplhost_tables_node_table(smplhost_tables* h, double* out, int max_nodes)
{
    if (!h) return -1;
    const smplgpu_robot_desc* d = h->t.desc();
    if (d->n_nodes > max_nodes) return -1;
    std::vector<int> tree_of(d->n_nodes, -1);
    for (int t = 0; t < d->n_trees; ++t) {
        std::vector<int> st(1, d->tree_root[t]);
        while (!st.empty()) {
            int n = st.back();
            st.pop_back();
            tree_of[n] = t;
            if (d->node_left[n] >= 0) {
                st.push_back(d->node_left[n]);
                st.push_back(d->node_right[n]);
            }
        }
    }
    // first node of each tree, to report children tree-relative like the oracle
    std::vector<int> first(d->n_trees, d->n_nodes);
    for (int n = 0; n < d->n_nodes; ++n) {
        if (tree_of[n] >= 0) first[tree_of[n]] = std::min(first[tree_of[n]], n);
    }
    for (int n = 0; n < d->n_nodes; ++n) {
        double* o = out + 8 * n;
        o[0] = d->node_center[3 * n]; o[1] = d->node_center[3 * n + 1]; o[2] = d->node_center[3 * n + 2];
        o[3] = d->node_radius[n];
        const int f = first[tree_of[n]];
        o[4] = d->node_left[n] < 0 ? -1 : d->node_left[n] - f;
        o[5] = d->node_right[n] < 0 ? -1 : d->node_right[n] - f;
        o[6] = tree_of[n];
        o[7] = d->node_link[n];
    }
    return d->n_nodes;
}

int smplhost_tables_motion_weights(smplhost_tables* h, double* weights, int32_t* types)
{
    if (!h) return -1;
    const smplgpu_robot_desc* d = h->t.desc();
    for (int i = 0; i < d->dof; ++i) {
        weights[i] = d->var_motion_weight[i];
        types[i] = d->var_type[i];
    }
    return 0;
}

int smplhost_tables_pairs(smplhost_tables* h, int32_t* out, int max_pairs)
{
    if (!h) return -1;
    const smplgpu_robot_desc* d = h->t.desc();
    for (int p = 0; p < d->n_pairs && p < max_pairs; ++p) {
        out[2 * p] = d->pair_a[p];
        out[2 * p + 1] = d->pair_b[p];
    }
    return d->n_pairs;
}

} // extern "C"
