#include "scene_ingest.h"

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

} // namespace smplhost
