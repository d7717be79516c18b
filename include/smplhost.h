/*
 * smplhost.h -- flat C entry points of the host-side C++ library
 * (smpl_b200/host): the model builder that produces the tables of smplgpu.h
 * from a robot description, and the adapter objects that mirror the reference's
 * plugin interfaces.  Used by the Python plumbing (tests, bench.py) via ctypes;
 * a C++ caller would use the classes in smpl_b200/host/ directly.
 */
#ifndef SMPLHOST_H
#define SMPLHOST_H

#include <stdint.h>

#include "smplgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct smplhost_tables smplhost_tables;

const char* smplhost_last_error(void);

/* ---- model builder (RobotCollisionModel::init + CollisionSpace::init, collision_space.cpp:649-739) ---- */
smplhost_tables* smplhost_tables_load(const char* robot_path);
void smplhost_tables_destroy(smplhost_tables* h);
int smplhost_tables_configure(smplhost_tables* h, const char* group, const char* planning_joints_csv);
/* CollisionSpace::setJointPosition (collision_space.cpp:166-183) for non-planning variables */
int smplhost_tables_set_joint(smplhost_tables* h, const char* variable, double value);
/* CollisionSpace::setAllowedCollisionMatrix with the description's acm records (call_planner.cpp:441-1527, 1631) */
int smplhost_tables_use_file_acm(smplhost_tables* h);
int smplhost_tables_set_acm_entry(smplhost_tables* h, const char* a, const char* b, int allowed);
/* AttachedBodiesCollisionModel::attachBody with a ready spheres model (attached_bodies_collision_model.cpp:70-141) */
int smplhost_tables_attach_spheres(smplhost_tables* h, const char* id, const char* link,
                                   const double* centers, int n, double radius);
int smplhost_tables_detach(smplhost_tables* h, const char* id);
/* KDLRobotModel::init + setPlanningLink + setKinematicsToPlanningTransform (kdl_robot_model.cpp:59-171) */
int smplhost_tables_set_planning_chain(smplhost_tables* h, const char* root, const char* tip,
                                       const char* planning_link, const double* T_kin_to_planning /*[12] or NULL*/,
                                       const double* xyz_offset /*[3] or NULL*/);
int smplhost_tables_dof(smplhost_tables* h);
int smplhost_tables_limits(smplhost_tables* h, double* mins, double* maxs, uint8_t* continuous);
const smplgpu_robot_desc* smplhost_tables_desc(smplhost_tables* h);
/* smplgpu_set_robot(ctx, desc) */
int smplhost_tables_apply(smplhost_tables* h, smplgpu_ctx* ctx);
/* world-frame voxels of out-of-group links (SelfCollisionModelImpl::updateGroup, self_collision_model.cpp:616-760) */
int smplhost_tables_outside_voxels(smplhost_tables* h, const double** xyz);
/* inspection, for parity tests against the oracle's tables */
int smplhost_tables_node_table(smplhost_tables* h, double* out /*[n][8]*/, int max_nodes);
int smplhost_tables_motion_weights(smplhost_tables* h, double* weights, int32_t* types);
int smplhost_tables_pairs(smplhost_tables* h, int32_t* out /*[n][2]*/, int max_pairs);

#ifdef __cplusplus
}
#endif

#endif /* SMPLHOST_H */
