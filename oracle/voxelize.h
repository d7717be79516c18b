// ORACLE (test infrastructure only -- never linked into the product path).
//
// CPU restatement of the reference's scene ingest: shapes -> triangle mesh ->
// surface voxels -> OccupancyGrid::addPointsToField.
//   CreateIndexedBoxMesh / SphereMesh / CylinderMesh / ConeMesh   smpl/src/geometry/mesh_utils.cpp:39-300
//   VoxelizeBox / VoxelizeMesh (pose, res, voxel_origin)   smpl/src/geometry/voxelize.cpp:673-736, 962-1054
//   VoxelizeMeshAwesome       voxelize.cpp:134-146
//   VoxelizeTriangle          smpl/include/smpl/geometry/detail/voxelize.hpp:45-181
//   Distance (capsule)        voxelize.cpp:626-649
//   ScanFill                  voxelize.cpp:561-606 (fill = true; no caller of the collision model passes it)
//   ExtractVoxels             voxelize.cpp:206-222
//   ComputeAxisAlignedBoundingBox  voxelize.cpp:224-259
//   PivotDiscretizer / HalfResDiscretizer   smpl/include/smpl/geometry/discretize.h:41-112
//   VoxelGrid                 smpl/include/smpl/geometry/voxel_grid.h:119-147 (extent), :343-384 (conversions)
//   callers: sbpl_collision_checking/src/voxel_operations.cpp:357-369 (fill = false, voxel origin = grid origin),
//            world_collision_model.cpp:193-234, attached_bodies_collision_model.cpp:264-309 (origin 0)
//
// Eigen arithmetic is restated as in omath.h (Eigen 3.3 on x86-64: 3-term sums left to right, normalize() divides).
// Checked against the reference's own voxelize.cpp compiled with a stand-in Eigen of the same operation order
// (oracle/ref_voxelize_shim.cpp, tests/test_oracle_voxelize.py): that pins the control flow (which cells are
// visited, which edges are tested -- the reference never tests the edge p1-p2 and tests p1-p3 twice), not Eigen.
#ifndef ORACLE_VOXELIZE_H
#define ORACLE_VOXELIZE_H

#include <vector>

#include "omath.h"

namespace oracle {

/// mesh_utils.cpp:39-113: 8 vertices, 12 triangles (36 indices), appended
void CreateIndexedBoxMesh(double length, double width, double height, std::vector<Vec3>& vertices, std::vector<int>& indices);

/// mesh_utils.cpp:116-205 (VoxelizeSphere passes 7 lines of longitude, 8 of latitude: voxelize.cpp:741-811),
/// :208-264 (16 rim points), :268-300 (16 rim points); all appended, indices relative to this shape's first vertex
void CreateIndexedSphereMesh(double radius, int longitude_count, int latitude_count, std::vector<Vec3>& vertices, std::vector<int>& indices);
void CreateIndexedCylinderMesh(double radius, double length, std::vector<Vec3>& vertices, std::vector<int>& indices);
void CreateIndexedConeMesh(double radius, double height, std::vector<Vec3>& vertices, std::vector<int>& indices);

/// One of the reference's two voxel grids: cells centred on pivot + i res (PivotVoxelGrid) or on (i + 1/2) res
/// (HalfResVoxelGrid)
struct VoxelDiscretizer
{
    bool half_res;
    double res;
    double pivot[3];
    int discretize(int axis, double d) const;
    double continuize(int axis, int i) const;
};

/// geometry::VoxelizeMesh(vertices, triangles, res, voxel_origin, voxels, fill) (voxel_origin == nullptr: the
/// half-res overload): voxel centres in ExtractVoxels order, appended
void VoxelizeMesh(const std::vector<Vec3>& vertices, const std::vector<int>& triangles, double res,
                  const double* voxel_origin, bool fill, std::vector<Vec3>& voxels);

/// geometry::VoxelizeBox(length, width, height, pose, res, voxel_origin, voxels, fill)
void VoxelizeBox(double length, double width, double height, const Affine3& pose, double res,
                 const double* voxel_origin, bool fill, std::vector<Vec3>& voxels);

} // namespace oracle

#endif
