// Scene ingest, host side: collision shapes -> triangle meshes in the grid frame.  The
// meshes are voxelised on the device (smplgpu_voxelize_mesh / smplgpu_build_distance_field_from_meshes).
//   geometry::CreateIndexedBoxMesh      smpl/src/geometry/mesh_utils.cpp:39-113
//   TransformVertices                   smpl/src/geometry/voxelize.cpp:608-615 (pose * vertex, Eigen Affine3d)
//   VoxelizeBox(length, width, height, pose, ...)   voxelize.cpp:690-736
//   GetCollisionCube / env file         smpl_test/src/call_planner.cpp (id x y z dx dy dz, identity orientation)
#ifndef SMPLHOST_SCENE_INGEST_H
#define SMPLHOST_SCENE_INGEST_H

#include <cstdint>
#include <vector>

namespace smplhost {

/// Appends the 8 vertices (transformed by pose, a 3x4 row-major rigid transform) and 12 triangles of a box; triangle
/// indices refer to the concatenated vertex array.
void AppendBoxMesh(double length, double width, double height, const double* pose3x4,
                   std::vector<double>& vertices, std::vector<int32_t>& triangles);

} // namespace smplhost

#endif
