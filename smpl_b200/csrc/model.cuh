// Device-resident robot / grid tables shared by the kernels.
#pragma once

#include <stdint.h>

namespace smplgpu {

constexpr int MAX_DOF = 16;
constexpr int MAX_LINKS = 48;
constexpr int MAX_NODES = 2048;
constexpr int MAX_TREES = 64;
constexpr int MAX_PAIRS = 1024;
constexpr int MAX_ALLOWED = 256;
constexpr int MAX_SEGMENTS = 32;
constexpr int MAX_TREE_DEPTH = 48;   // DFS stack bound (checked on the host)

// Robot tables as the kernels read them (one copy in global memory, hot in L1).
struct DevModel
{
    int dof, n_links, n_nodes, n_trees, n_pairs, n_allowed, n_slots, n_segments;
    int n_robot_trees, n_robot_pairs;   // trees of robot links come first; pairs with both trees among them

    // links, topological order
    int link_parent[MAX_LINKS];
    int link_joint[MAX_LINKS];
    int link_var[MAX_LINKS];
    int link_slot[MAX_LINKS];        // shared-memory slot holding T_link, -1 = not kept
    int link_tree_begin[MAX_LINKS];  // range into tree_by_link[]: trees rooted on this link
    int link_tree_end[MAX_LINKS];
    double link_const[MAX_LINKS];
    double link_origin[MAX_LINKS][12];
    double link_axis[MAX_LINKS][3];
    double link_base[MAX_LINKS][12];

    // sphere tree nodes
    int node_link[MAX_NODES];
    int node_left[MAX_NODES];
    int node_right[MAX_NODES];
    int node_thresh[MAX_NODES];      // sphere ok  <=>  d2(cell) >= node_thresh
    double node_center[MAX_NODES][3];
    double node_radius[MAX_NODES];

    int tree_root[MAX_TREES];
    int tree_by_link[MAX_TREES];     // tree indices sorted by link
    int pair_a[MAX_PAIRS];
    int pair_b[MAX_PAIRS];
    int allowed_a[MAX_ALLOWED];
    int allowed_b[MAX_ALLOWED];

    int var_type[MAX_DOF];
    double var_weight[MAX_DOF];
    double var_min[MAX_DOF];
    double var_max[MAX_DOF];
    double var_min_norm[MAX_DOF];    // angles::normalize_angle(var_min)

    // planning-link chain
    int seg_kind[MAX_SEGMENTS];
    int seg_var[MAX_SEGMENTS];
    double seg_axis[MAX_SEGMENTS][3];
    double seg_origin[MAX_SEGMENTS][3];
    double seg_f_tip[MAX_SEGMENTS][12];
    double T_kin_to_planning[12];
    double xyz_offset[3];
    int has_offset;
};

// Distance field read parameters (DistanceMap::worldToGrid, distance_map.hpp:520-527)
struct GridParams
{
    int nx, ny, nz;
    double inv_res;
    double ox, oy, oz;   // origin - res, per axis
};

} // namespace smplgpu
