// ORACLE (test infrastructure only -- never linked into the product path).
//
// Minimal double-precision vector / affine helpers standing in for Eigen3
// (un-vendored third-party dependency of the reference, version unpinned --
// SURVEY.md section 8c).  Operation order follows Eigen's fixed-size
// coefficient-based products as used by Affine3d:
//   (A*B).linear      = A.linear * B.linear          dot3 = (a0*b0 + a1*b1) + a2*b2
//   (A*B).translation = A.linear * B.translation + A.translation
//   A*v               = A.linear * v + A.translation
// No FMA contraction: the reference builds with plain -O2 for x86-64
// (smpl/CMakeLists.txt:7-8,44), so oracle/Makefile passes -ffp-contract=off.
#ifndef ORACLE_OMATH_H
#define ORACLE_OMATH_H

#include <cmath>

namespace oracle {

struct Vec3
{
    double x, y, z;
    Vec3() : x(0.0), y(0.0), z(0.0) { }
    Vec3(double x_, double y_, double z_) : x(x_), y(y_), z(z_) { }
    double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};

inline Vec3 operator+(const Vec3& a, const Vec3& b) { return Vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline Vec3 operator-(const Vec3& a, const Vec3& b) { return Vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline Vec3 operator*(double s, const Vec3& a) { return Vec3(s * a.x, s * a.y, s * a.z); }
inline Vec3 operator*(const Vec3& a, double s) { return Vec3(a.x * s, a.y * s, a.z * s); }
inline Vec3 operator/(const Vec3& a, double s) { return Vec3(a.x / s, a.y / s, a.z / s); }
inline double dot(const Vec3& a, const Vec3& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline double squaredNorm(const Vec3& a) { return dot(a, a); }
inline double norm(const Vec3& a) { return std::sqrt(squaredNorm(a)); }
inline Vec3 normalized(const Vec3& a)
{
    // Eigen: n = norm(); n > 0 ? a / n : a
    double n = norm(a);
    if (n > 0.0) {
        return a / n;
    }
    return a;
}

/// 3x4 affine transform, m[r][c], c == 3 is the translation column.
struct Affine3
{
    double m[3][4];

    static Affine3 Identity()
    {
        Affine3 t;
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 4; ++c) {
                t.m[r][c] = (r == c) ? 1.0 : 0.0;
            }
        }
        return t;
    }

    double operator()(int r, int c) const { return m[r][c]; }
    double& operator()(int r, int c) { return m[r][c]; }
    Vec3 translation() const { return Vec3(m[0][3], m[1][3], m[2][3]); }
};

inline Affine3 operator*(const Affine3& a, const Affine3& b)
{
    Affine3 r;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) {
            r.m[i][j] = (a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j]) + a.m[i][2] * b.m[2][j];
        }
        r.m[i][3] = ((a.m[i][0] * b.m[0][3] + a.m[i][1] * b.m[1][3]) + a.m[i][2] * b.m[2][3]) + a.m[i][3];
    }
    return r;
}

inline Vec3 operator*(const Affine3& a, const Vec3& v)
{
    return Vec3(
        ((a.m[0][0] * v.x + a.m[0][1] * v.y) + a.m[0][2] * v.z) + a.m[0][3],
        ((a.m[1][0] * v.x + a.m[1][1] * v.y) + a.m[1][2] * v.z) + a.m[1][3],
        ((a.m[2][0] * v.x + a.m[2][1] * v.y) + a.m[2][2] * v.z) + a.m[2][3]);
}

/// Eigen::AngleAxisd(angle, axis).toRotationMatrix() (Eigen/src/Geometry/AngleAxis.h)
inline Affine3 AngleAxis(double angle, const Vec3& axis)
{
    Affine3 r = Affine3::Identity();
    const double s = std::sin(angle);
    const double c = std::cos(angle);
    Vec3 sin_axis = s * axis;
    Vec3 cos1_axis = (1.0 - c) * axis;
    double tmp;
    tmp = cos1_axis.x * axis.y;
    r.m[0][1] = tmp - sin_axis.z;
    r.m[1][0] = tmp + sin_axis.z;
    tmp = cos1_axis.x * axis.z;
    r.m[0][2] = tmp + sin_axis.y;
    r.m[2][0] = tmp - sin_axis.y;
    tmp = cos1_axis.y * axis.z;
    r.m[1][2] = tmp - sin_axis.x;
    r.m[2][1] = tmp + sin_axis.x;
    r.m[0][0] = cos1_axis.x * axis.x + c;
    r.m[1][1] = cos1_axis.y * axis.y + c;
    r.m[2][2] = cos1_axis.z * axis.z + c;
    return r;
}

inline Affine3 Translation(double x, double y, double z)
{
    Affine3 r = Affine3::Identity();
    r.m[0][3] = x;
    r.m[1][3] = y;
    r.m[2][3] = z;
    return r;
}

/// urdf::Rotation::setFromRPY followed by Eigen::Quaterniond::toRotationMatrix,
/// i.e. what poseUrdfToEigen does for a joint origin (robot_collision_model.cpp:362-366).
inline Affine3 FromXyzRpy(double x, double y, double z, double roll, double pitch, double yaw)
{
    double phi = roll / 2.0, the = pitch / 2.0, psi = yaw / 2.0;
    double qx = std::sin(phi) * std::cos(the) * std::cos(psi) - std::cos(phi) * std::sin(the) * std::sin(psi);
    double qy = std::cos(phi) * std::sin(the) * std::cos(psi) + std::sin(phi) * std::cos(the) * std::sin(psi);
    double qz = std::cos(phi) * std::cos(the) * std::sin(psi) - std::sin(phi) * std::sin(the) * std::cos(psi);
    double qw = std::cos(phi) * std::cos(the) * std::cos(psi) + std::sin(phi) * std::sin(the) * std::sin(psi);
    // urdf normalises the quaternion
    double s = std::sqrt(qx * qx + qy * qy + qz * qz + qw * qw);
    if (std::fabs(s) < 1e-5) {
        qx = 0.0; qy = 0.0; qz = 0.0; qw = 1.0;
    } else {
        qx /= s; qy /= s; qz /= s; qw /= s;
    }
    const double tx = 2.0 * qx, ty = 2.0 * qy, tz = 2.0 * qz;
    const double twx = tx * qw, twy = ty * qw, twz = tz * qw;
    const double txx = tx * qx, txy = ty * qx, txz = tz * qx;
    const double tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
    Affine3 r = Affine3::Identity();
    r.m[0][0] = 1.0 - (tyy + tzz);
    r.m[0][1] = txy - twz;
    r.m[0][2] = txz + twy;
    r.m[1][0] = txy + twz;
    r.m[1][1] = 1.0 - (txx + tzz);
    r.m[1][2] = tyz - twx;
    r.m[2][0] = txz - twy;
    r.m[2][1] = tyz + twx;
    r.m[2][2] = 1.0 - (txx + tyy);
    r.m[0][3] = x;
    r.m[1][3] = y;
    r.m[2][3] = z;
    return r;
}

} // namespace oracle

#endif
