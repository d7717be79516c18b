// ORACLE (test infrastructure only).  See voxelize.h.
#include "voxelize.h"

#include <algorithm>
#include <cmath>

namespace oracle {

void CreateIndexedBoxMesh(double length, double width, double height, std::vector<Vec3>& vertices, std::vector<int>& indices)
{
    const double hx = 0.5 * length, hy = 0.5 * width, hz = 0.5 * height;
    // lbb, rbb, ltb, rtb, lbf, rbf, ltf, rtf (l/r = -/+ x, b/t = -/+ y, b/f = -/+ z)
    const double sx[8] = { -1, 1, -1, 1, -1, 1, -1, 1 };
    const double sy[8] = { -1, -1, 1, 1, -1, -1, 1, 1 };
    const double sz[8] = { -1, -1, -1, -1, 1, 1, 1, 1 };
    const int base = (int)vertices.size();
    for (int i = 0; i < 8; ++i) {
        vertices.push_back(Vec3(sx[i] * hx, sy[i] * hy, sz[i] * hz));
    }
    static const int tri[36] = {
        0, 2, 1, 1, 2, 3,   // back
        5, 1, 7, 1, 3, 7,   // right
        5, 7, 4, 4, 7, 6,   // front
        4, 6, 0, 0, 6, 2,   // left
        6, 7, 2, 7, 3, 2,   // top
        5, 0, 1, 4, 5, 0 }; // bottom
    for (int i = 0; i < 36; ++i) {
        indices.push_back(tri[i]);   // the reference's indices are NOT offset by earlier vertices either
    }
    (void)base;
}

void CreateIndexedSphereMesh(double radius, int longitude_count, int latitude_count, std::vector<Vec3>& vertices, std::vector<int>& indices)
{
    vertices.push_back(Vec3(0.0, 0.0, radius));   // north pole
    const double theta_inc = M_PI / (latitude_count + 1);
    const double phi_inc = (2.0 * M_PI) / longitude_count;
    for (int t = 0; t < latitude_count; ++t) {
        for (int p = 0; p < longitude_count; ++p) {
            const double theta = (t + 1) * theta_inc, phi = p * phi_inc;
            vertices.push_back(Vec3(radius * std::sin(theta) * std::cos(phi), radius * std::sin(theta) * std::sin(phi),
                                    radius * std::cos(theta)));
        }
    }
    vertices.push_back(Vec3(0.0, 0.0, -radius));  // south pole
    for (int i = 0; i < longitude_count; ++i) {   // fan around the north pole
        indices.push_back(0);
        indices.push_back(i + 1);
        indices.push_back(i == longitude_count - 1 ? 1 : i + 2);
    }
    for (int i = 0; i < latitude_count - 1; ++i) {   // quads between neighbouring lines of latitude
        for (int j = 0; j < longitude_count; ++j) {
            const int base = i * longitude_count + j + 1;
            const int below = base + longitude_count;
            int below_right = below + 1, right = base + 1;
            if ((below_right - 1) / longitude_count != (below - 1) / longitude_count) below_right -= longitude_count;
            if ((right - 1) / longitude_count != (base - 1) / longitude_count) right -= longitude_count;
            const int quad[6] = { base, below, below_right, base, below_right, right };
            indices.insert(indices.end(), quad, quad + 6);
        }
    }
    const int last = (int)vertices.size() - 1;    // the reference uses the size of the WHOLE vertex array here
    for (int i = 0; i < longitude_count; ++i) {   // fan around the south pole
        indices.push_back(last);
        indices.push_back(i == 0 ? last - longitude_count : last - i);
        indices.push_back(last - (i + 1));
    }
}

void CreateIndexedCylinderMesh(double radius, double length, std::vector<Vec3>& vertices, std::vector<int>& indices)
{
    const int rim = 16;
    for (int cap = 0; cap < 2; ++cap) {
        for (int i = 0; i < rim; ++i) {
            const double theta = 2.0 * M_PI * (double)i / double(rim);
            vertices.push_back(Vec3(radius * std::cos(theta), radius * std::sin(theta), cap == 0 ? 0.5 * length : -0.5 * length));
        }
    }
    vertices.push_back(Vec3(0.0, 0.0, 0.5 * length));
    vertices.push_back(Vec3(0.0, 0.0, -0.5 * length));
    for (int i = 0; i < rim; ++i) {
        const int n = (i + 1) % rim;
        const int side[6] = { i, n, i + rim, n, n + rim, i + rim };
        indices.insert(indices.end(), side, side + 6);
    }
    for (int i = 0; i < rim; ++i) {
        const int top[3] = { 2 * rim, (i + 1) % rim, i };
        indices.insert(indices.end(), top, top + 3);
    }
    for (int i = 0; i < rim; ++i) {
        const int bottom[3] = { 2 * rim + 1, (i + 1) % rim + rim, i + rim };
        indices.insert(indices.end(), bottom, bottom + 3);
    }
}

void CreateIndexedConeMesh(double radius, double height, std::vector<Vec3>& vertices, std::vector<int>& indices)
{
    const int rim = 16;
    for (int i = 0; i < rim; ++i) {
        const double theta = 2.0 * M_PI * (double)i / (double)rim;
        vertices.push_back(Vec3(radius * std::cos(theta), radius * std::sin(theta), -0.5 * height));
    }
    vertices.push_back(Vec3(0.0, 0.0, 0.5 * height));
    vertices.push_back(Vec3(0.0, 0.0, -0.5 * height));
    for (int apex = rim; apex <= rim + 1; ++apex) {   // the mantle, then the bottom fan
        for (int i = 0; i < rim; ++i) {
            const int tri[3] = { i, (i + 1) % rim, apex };
            indices.insert(indices.end(), tri, tri + 3);
        }
    }
}

int VoxelDiscretizer::discretize(int axis, double d) const
{
    if (half_res) {
        return (d >= 0) ? (int)(d / res) : ((int)(d / res) - 1);
    }
    return (int)std::floor((d - pivot[axis]) / res + 0.5);
}

double VoxelDiscretizer::continuize(int axis, int i) const
{
    if (half_res) {
        return (double)i * res + 0.5 * res;
    }
    return pivot[axis] + i * res;
}

namespace {

struct Grid
{
    VoxelDiscretizer disc;
    int min_gc[3], max_gc[3];
    std::vector<unsigned char> cells;
    int size(int a) const { return max_gc[a] - min_gc[a] + 1; }
    size_t index(int gx, int gy, int gz) const
    {
        return ((size_t)(gx - min_gc[0]) * size(1) + (size_t)(gy - min_gc[1])) * size(2) + (size_t)(gz - min_gc[2]);
    }
};

Vec3 cross(const Vec3& a, const Vec3& b)
{
    return Vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

Vec3 neg(const Vec3& a) { return Vec3(-a.x, -a.y, -a.z); }

/// voxelize.cpp:626-649: squared distance of x to the segment p-q when its projection falls on the segment
/// and the distance is within the radius, else -1
double CapsuleDistance(const Vec3& p, const Vec3& q, double radius_sqrd, const Vec3& x)
{
    const Vec3 pq = q - p;
    const Vec3 px = x - p;
    const double d = dot(px, pq);
    if (d < 0.0 || d > squaredNorm(pq)) {
        return -1.0;
    }
    const double dsq = squaredNorm(px) - (d * d) / squaredNorm(pq);
    if (dsq > radius_sqrd) {
        return -1.0;
    }
    return dsq;
}

double sign(double v) { return (v == 0) ? 0.0 : ((v > 0) ? 1.0 : -1.0); }

/// detail/voxelize.hpp:45-181
void VoxelizeTriangle(const Vec3& a, const Vec3& b, const Vec3& c, Grid& vg)
{
    Vec3 p1 = a, p2 = b, p3 = c;
    const double det = norm(cross(p2 - p1, p3 - p1));
    if (det == 0) {
        return;
    }
    if (det < 0.0) {   // never true for a norm; kept for the shape of the reference
        std::swap(p1, p3);
    }
    const double rc = std::sqrt(3.0) * 0.5 * vg.disc.res;
    const double rc2 = rc * rc;
    const Vec3 u = p2 - p1, v = p3 - p2, w = p1 - p3;
    const Vec3 n = normalized(cross(u, v));
    const double k = 0.5774;
    double ca = dot(Vec3(-k, -k, -k), n);
    for (int i = 1; i < 8; ++i) {
        ca = std::max(ca, dot(Vec3((i & 4) ? k : -k, (i & 2) ? k : -k, (i & 1) ? k : -k), n));
    }
    const double t = rc * ca;
    const double d = -dot(n, p1);
    const Vec3 e1 = normalized(neg(cross(u, n)));
    const Vec3 e2 = normalized(neg(cross(v, n)));
    const Vec3 e3 = normalized(neg(cross(w, n)));
    const double d1 = -dot(e1, p1), d2 = -dot(e2, p2), d3 = -dot(e3, p3);

    const double lo[3] = { std::min(a.x, std::min(b.x, c.x)), std::min(a.y, std::min(b.y, c.y)), std::min(a.z, std::min(b.z, c.z)) };
    const double hi[3] = { std::max(a.x, std::max(b.x, c.x)), std::max(a.y, std::max(b.y, c.y)), std::max(a.z, std::max(b.z, c.z)) };
    int mn[3], mx[3];
    for (int ax = 0; ax < 3; ++ax) {
        mn[ax] = vg.disc.discretize(ax, lo[ax]);
        mx[ax] = vg.disc.discretize(ax, hi[ax]);
    }
    for (int gx = mn[0]; gx <= mx[0]; ++gx) {
    for (int gy = mn[1]; gy <= mx[1]; ++gy) {
    for (int gz = mn[2]; gz <= mx[2]; ++gz) {
        unsigned char& cell = vg.cells[vg.index(gx, gy, gz)];
        if (cell) {
            continue;
        }
        const Vec3 p(vg.disc.continuize(0, gx), vg.disc.continuize(1, gy), vg.disc.continuize(2, gz));
        if (squaredNorm(p - p1) <= rc2 || squaredNorm(p - p2) <= rc2 || squaredNorm(p - p3) <= rc2) {
            cell = 1;   // a vertex fills the voxel
        } else if (CapsuleDistance(p1, p3, rc2, p) != -1.0 || CapsuleDistance(p2, p3, rc2, p) != -1.0 ||
                   CapsuleDistance(p3, p1, rc2, p) != -1.0) {
            cell = 1;   // an edge fills the voxel (p1-p2 is never asked, p1-p3 twice: voxelize.hpp:146-148)
        } else if (sign(dot(n, p) + (d + t)) != sign(dot(n, p) + (d - t)) &&
                   dot(e1, p) + d1 > 0.0 && dot(e2, p) + d2 > 0.0 && dot(e3, p) + d3 > 0.0) {
            cell = 1;   // within the slab around the plane and inside the three edge planes
        }
    }
    }
    }
}

/// voxelize.cpp:561-606
void ScanFill(Grid& vg)
{
    const int nx = vg.size(0), ny = vg.size(1), nz = vg.size(2);
    enum { OUTSIDE = 0, ON_BOUNDARY_FROM_OUTSIDE = 1, INSIDE = 2, ON_BOUNDARY_FROM_INSIDE = 4 };
    for (int x = 0; x < nx; ++x) {
    for (int y = 0; y < ny; ++y) {
        unsigned char* col = &vg.cells[((size_t)x * ny + y) * nz];
        int state = OUTSIDE;
        for (int z = 0; z < nz; ++z) {
            if (state == OUTSIDE && col[z]) {
                state = ON_BOUNDARY_FROM_OUTSIDE;
            } else if (state == ON_BOUNDARY_FROM_OUTSIDE && !col[z]) {
                bool all_empty = true;
                for (int l = z; l < nz; ++l) {
                    all_empty &= !col[l];
                }
                if (all_empty) {
                    state = OUTSIDE;
                } else {
                    state = INSIDE;
                    col[z] = 1;
                }
            } else if (state == INSIDE && !col[z]) {
                col[z] = 1;
            } else if (state == INSIDE && col[z]) {
                state = ON_BOUNDARY_FROM_INSIDE;
            } else if (state == ON_BOUNDARY_FROM_INSIDE && !col[z]) {
                state = OUTSIDE;
            }
        }
    }
    }
}

} // namespace

void VoxelizeMesh(const std::vector<Vec3>& vertices, const std::vector<int>& triangles, double res,
                  const double* voxel_origin, bool fill, std::vector<Vec3>& voxels)
{
    if (triangles.size() % 3 != 0 || vertices.empty()) {
        return;
    }
    Vec3 mn = vertices[0], mx = vertices[0];
    for (const Vec3& p : vertices) {
        mn = Vec3(std::min(mn.x, p.x), std::min(mn.y, p.y), std::min(mn.z, p.z));
        mx = Vec3(std::max(mx.x, p.x), std::max(mx.y, p.y), std::max(mx.z, p.z));
    }
    const Vec3 size = mx - mn;
    Grid vg;
    vg.disc.half_res = voxel_origin == nullptr;
    vg.disc.res = res;
    for (int a = 0; a < 3; ++a) {
        vg.disc.pivot[a] = voxel_origin ? voxel_origin[a] : 0.0;
        vg.min_gc[a] = vg.disc.discretize(a, mn[a]);
        vg.max_gc[a] = vg.disc.discretize(a, mn[a] + size[a]);   // voxel_grid.h:139-141
    }
    vg.cells.assign((size_t)vg.size(0) * vg.size(1) * vg.size(2), 0);
    for (size_t i = 0; i + 2 < triangles.size(); i += 3) {
        VoxelizeTriangle(vertices[triangles[i]], vertices[triangles[i + 1]], vertices[triangles[i + 2]], vg);
    }
    if (fill) {
        ScanFill(vg);
    }
    for (int gx = vg.min_gc[0]; gx <= vg.max_gc[0]; ++gx) {
    for (int gy = vg.min_gc[1]; gy <= vg.max_gc[1]; ++gy) {
    for (int gz = vg.min_gc[2]; gz <= vg.max_gc[2]; ++gz) {
        if (vg.cells[vg.index(gx, gy, gz)]) {
            voxels.push_back(Vec3(vg.disc.continuize(0, gx), vg.disc.continuize(1, gy), vg.disc.continuize(2, gz)));
        }
    }
    }
    }
}

void VoxelizeBox(double length, double width, double height, const Affine3& pose, double res,
                 const double* voxel_origin, bool fill, std::vector<Vec3>& voxels)
{
    std::vector<Vec3> vertices;
    std::vector<int> triangles;
    CreateIndexedBoxMesh(length, width, height, vertices, triangles);
    for (Vec3& p : vertices) {
        p = pose * p;   // TransformVertices
    }
    VoxelizeMesh(vertices, triangles, res, voxel_origin, fill, voxels);
}

} // namespace oracle
