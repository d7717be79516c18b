// ORACLE (test infrastructure only).  C shim around the REFERENCE's own voxeliser (smpl/src/geometry/voxelize.cpp
// + mesh_utils.cpp, compiled where they lie by `make -C oracle ref` against oracle/ref_stubs/eigen_arith -- see the
// header of that stand-in for what this does and does not pin) -> oracle/_ref/libref_voxelize.so.
#include <cstdint>
#include <vector>

#include <Eigen/Dense>
#include <smpl/geometry/mesh_utils.h>
#include <smpl/geometry/voxelize.h>

namespace {

Eigen::Affine3d ToAffine(const double* m /*3x4 row-major*/)
{
    Eigen::Affine3d t;
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) t.linear()(r, c) = m[4 * r + c];
        t.translation()[r] = m[4 * r + 3];
    }
    return t;
}

int Emit(const std::vector<Eigen::Vector3d>& voxels, double* out, int max_out)
{
    if ((int)voxels.size() > max_out) {
        return -(int)voxels.size();
    }
    for (size_t i = 0; i < voxels.size(); ++i) {
        out[3 * i] = voxels[i].x();
        out[3 * i + 1] = voxels[i].y();
        out[3 * i + 2] = voxels[i].z();
    }
    return (int)voxels.size();
}

} // namespace

extern "C" {

/// geometry::VoxelizeMesh(vertices, triangles, res[, voxel_origin], voxels, fill); returns the voxel count
/// (negative count when max_out is too small)
int ref_voxelize_mesh(const double* vertices, int nv, const int32_t* triangles, int nt, double res,
                      const double* voxel_origin, int fill, double* out, int max_out)
{
    std::vector<Eigen::Vector3d> v(nv);
    for (int i = 0; i < nv; ++i) v[i] = Eigen::Vector3d(vertices[3 * i], vertices[3 * i + 1], vertices[3 * i + 2]);
    std::vector<int> t(triangles, triangles + 3 * (size_t)nt);
    std::vector<Eigen::Vector3d> voxels;
    if (voxel_origin) {
        sbpl::geometry::VoxelizeMesh(v, t, res, Eigen::Vector3d(voxel_origin[0], voxel_origin[1], voxel_origin[2]), voxels, fill != 0);
    } else {
        sbpl::geometry::VoxelizeMesh(v, t, res, voxels, fill != 0);
    }
    return Emit(voxels, out, max_out);
}

/// geometry::VoxelizeBox(length, width, height, pose, res[, voxel_origin], voxels, fill)
int ref_voxelize_box(double length, double width, double height, const double* pose3x4, double res,
                     const double* voxel_origin, int fill, double* out, int max_out)
{
    std::vector<Eigen::Vector3d> voxels;
    const Eigen::Affine3d pose = ToAffine(pose3x4);
    if (voxel_origin) {
        sbpl::geometry::VoxelizeBox(length, width, height, pose, res,
                                    Eigen::Vector3d(voxel_origin[0], voxel_origin[1], voxel_origin[2]), voxels, fill != 0);
    } else {
        sbpl::geometry::VoxelizeBox(length, width, height, pose, res, voxels, fill != 0);
    }
    return Emit(voxels, out, max_out);
}

/// geometry::CreateIndexedBoxMesh: vertices[8][3], indices[36]
void ref_box_mesh(double length, double width, double height, double* vertices, int32_t* indices)
{
    std::vector<Eigen::Vector3d> v;
    std::vector<int> t;
    sbpl::geometry::CreateIndexedBoxMesh(length, width, height, v, t);
    for (size_t i = 0; i < v.size(); ++i) {
        vertices[3 * i] = v[i].x();
        vertices[3 * i + 1] = v[i].y();
        vertices[3 * i + 2] = v[i].z();
    }
    for (size_t i = 0; i < t.size(); ++i) indices[i] = t[i];
}

/// geometry::CreateIndexed{Box,Sphere,Cylinder,Cone}Mesh (kind 0..3) with the parameters VoxelizeSphere etc. pass
int ref_shape_mesh(int kind, const double* dims, double* vertices, int32_t* indices, int* n_indices)
{
    std::vector<Eigen::Vector3d> v;
    std::vector<int> t;
    if (kind == 0) sbpl::geometry::CreateIndexedBoxMesh(dims[0], dims[1], dims[2], v, t);
    else if (kind == 1) sbpl::geometry::CreateIndexedSphereMesh(dims[0], 7, 8, v, t);
    else if (kind == 2) sbpl::geometry::CreateIndexedCylinderMesh(dims[0], dims[1], v, t);
    else sbpl::geometry::CreateIndexedConeMesh(dims[0], dims[1], v, t);
    for (size_t i = 0; i < v.size(); ++i) {
        vertices[3 * i] = v[i].x();
        vertices[3 * i + 1] = v[i].y();
        vertices[3 * i + 2] = v[i].z();
    }
    for (size_t i = 0; i < t.size(); ++i) indices[i] = t[i];
    *n_indices = (int)t.size();
    return (int)v.size();
}

} // extern "C"
