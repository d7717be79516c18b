#include "scene_ingest.h"

#include <cmath>

namespace smplhost {

void AppendBoxMesh(double length, double width, double height, const double* pose3x4,
                   std::vector<double>& vertices, std::vector<int32_t>& triangles)
{
    const int32_t base = (int32_t)(vertices.size() / 3);
    const double h[3] = { 0.5 * length, 0.5 * width, 0.5 * height };
    // corner i: bit 0 = +x, bit 1 = +y, bit 2 = +z  (the reference's lbb, rbb, ltb, rtb, lbf, rbf, ltf, rtf)
    for (int i = 0; i < 8; ++i) {
        const double c[3] = { (i & 1) ? h[0] : -h[0], (i & 2) ? h[1] : -h[1], (i & 4) ? h[2] : -h[2] };
        for (int r = 0; r < 3; ++r) {
            // Affine3d * Vector3d: linear * v (3-term sum left to right) + translation
            const double* m = pose3x4 + 4 * r;
            vertices.push_back(((m[0] * c[0] + m[1] * c[1]) + m[2] * c[2]) + m[3]);
        }
    }
    static const int32_t faces[12][3] = {
        { 0, 2, 1 }, { 1, 2, 3 },     // back   (z-)
        { 5, 1, 7 }, { 1, 3, 7 },     // right  (x+)
        { 5, 7, 4 }, { 4, 7, 6 },     // front  (z+)
        { 4, 6, 0 }, { 0, 6, 2 },     // left   (x-)
        { 6, 7, 2 }, { 7, 3, 2 },     // top    (y+)
        { 5, 0, 1 }, { 4, 5, 0 } };   // bottom (y-)
    for (const auto& f : faces) {
        for (int k = 0; k < 3; ++k) triangles.push_back(base + f[k]);
    }
}

namespace {

const int kRim = 16, kLongitude = 7, kLatitude = 8;

// local (shape-frame) vertices and triangles -> appended in the grid frame
void emit(const std::vector<double>& local, const std::vector<int32_t>& tris, const double* pose3x4,
          std::vector<double>& vertices, std::vector<int32_t>& triangles)
{
    const int32_t base = (int32_t)(vertices.size() / 3);
    for (size_t i = 0; i + 2 < local.size(); i += 3) {
        for (int r = 0; r < 3; ++r) {
            const double* m = pose3x4 + 4 * r;
            vertices.push_back(((m[0] * local[i] + m[1] * local[i + 1]) + m[2] * local[i + 2]) + m[3]);
        }
    }
    for (int32_t t : tris) triangles.push_back(base + t);
}

} // namespace

void ShapeMeshSize(int kind, int* n_vertices, int* n_triangles)
{
    switch (kind) {
    case SHAPE_BOX: *n_vertices = 8; *n_triangles = 12; break;
    case SHAPE_SPHERE: *n_vertices = 2 + kLongitude * kLatitude; *n_triangles = 2 * kLongitude + 2 * (kLatitude - 1) * kLongitude; break;
    case SHAPE_CYLINDER: *n_vertices = 2 * kRim + 2; *n_triangles = 4 * kRim; break;
    case SHAPE_CONE: *n_vertices = kRim + 2; *n_triangles = 2 * kRim; break;
    default: *n_vertices = 0; *n_triangles = 0;
    }
}

bool AppendShapeMesh(int kind, const double* dims, const double* pose3x4,
                     std::vector<double>& vertices, std::vector<int32_t>& triangles)
{
    std::vector<double> v;
    std::vector<int32_t> t;
    auto vert = [&](double x, double y, double z) { v.push_back(x); v.push_back(y); v.push_back(z); };
    auto tri = [&](int a, int b, int c) { t.push_back(a); t.push_back(b); t.push_back(c); };
    if (kind == SHAPE_BOX) {
        AppendBoxMesh(dims[0], dims[1], dims[2], pose3x4, vertices, triangles);
        return true;
    }
    if (kind == SHAPE_SPHERE) {
        const double radius = dims[0];
        const int nlng = kLongitude, nlat = kLatitude;
        vert(0.0, 0.0, radius);
        const double theta_inc = M_PI / (nlat + 1), phi_inc = (2.0 * M_PI) / nlng;
        for (int a = 0; a < nlat; ++a) {
            const double theta = (a + 1) * theta_inc;
            for (int b = 0; b < nlng; ++b) {
                const double phi = b * phi_inc;
                vert(radius * std::sin(theta) * std::cos(phi), radius * std::sin(theta) * std::sin(phi), radius * std::cos(theta));
            }
        }
        vert(0.0, 0.0, -radius);
        auto ring = [&](int lat, int lng) { return 1 + lat * nlng + (lng % nlng); };   // vertex of a ring, wrapping around
        for (int b = 0; b < nlng; ++b) tri(0, ring(0, b), ring(0, b + 1));
        for (int a = 0; a + 1 < nlat; ++a) {
            for (int b = 0; b < nlng; ++b) {
                tri(ring(a, b), ring(a + 1, b), ring(a + 1, b + 1));
                tri(ring(a, b), ring(a + 1, b + 1), ring(a, b + 1));
            }
        }
        const int south = 1 + nlat * nlng;
        for (int b = 0; b < nlng; ++b) tri(south, b == 0 ? south - nlng : south - b, south - (b + 1));
    } else if (kind == SHAPE_CYLINDER) {
        const double radius = dims[0], length = dims[1];
        for (int cap = 0; cap < 2; ++cap) {
            for (int i = 0; i < kRim; ++i) {
                const double theta = 2.0 * M_PI * (double)i / double(kRim);
                vert(radius * std::cos(theta), radius * std::sin(theta), cap ? -0.5 * length : 0.5 * length);
            }
        }
        vert(0.0, 0.0, 0.5 * length);
        vert(0.0, 0.0, -0.5 * length);
        for (int i = 0; i < kRim; ++i) {
            const int n = (i + 1) % kRim;
            tri(i, n, i + kRim);
            tri(n, n + kRim, i + kRim);
        }
        for (int i = 0; i < kRim; ++i) tri(2 * kRim, (i + 1) % kRim, i);
        for (int i = 0; i < kRim; ++i) tri(2 * kRim + 1, (i + 1) % kRim + kRim, i + kRim);
    } else if (kind == SHAPE_CONE) {
        const double radius = dims[0], height = dims[1];
        for (int i = 0; i < kRim; ++i) {
            const double theta = 2.0 * M_PI * (double)i / (double)kRim;
            vert(radius * std::cos(theta), radius * std::sin(theta), -0.5 * height);
        }
        vert(0.0, 0.0, 0.5 * height);
        vert(0.0, 0.0, -0.5 * height);
        for (int i = 0; i < kRim; ++i) tri(i, (i + 1) % kRim, kRim);
        for (int i = 0; i < kRim; ++i) tri(i, (i + 1) % kRim, kRim + 1);
    } else {
        return false;
    }
    emit(v, t, pose3x4, vertices, triangles);
    return true;
}

} // namespace smplhost
