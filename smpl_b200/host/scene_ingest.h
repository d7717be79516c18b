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

/// The reference's other primitive meshes, built at the origin and moved by pose like AppendBoxMesh:
///   CreateIndexedSphereMesh(radius, 7, 8)   mesh_utils.cpp:116-205 (counts as VoxelizeSphere passes them, voxelize.cpp:741-811)
///   CreateIndexedCylinderMesh(radius, length), CreateIndexedConeMesh(radius, height)   mesh_utils.cpp:208-300
enum ShapeKind { SHAPE_BOX = 0, SHAPE_SPHERE = 1, SHAPE_CYLINDER = 2, SHAPE_CONE = 3 };

/// vertices / triangles a shape of this kind adds (box 8 / 12, sphere 58 / 112, cylinder 34 / 64, cone 18 / 32)
void ShapeMeshSize(int kind, int* n_vertices, int* n_triangles);

/// dims: box l, w, h; sphere r; cylinder r, length; cone r, height.  False for an unknown kind.
bool AppendShapeMesh(int kind, const double* dims, const double* pose3x4,
                     std::vector<double>& vertices, std::vector<int32_t>& triangles);

} // namespace smplhost

#endif
