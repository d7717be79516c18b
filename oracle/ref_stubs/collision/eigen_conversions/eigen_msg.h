// ORACLE (test infrastructure only).  Stand-in for an absent third-party header: just enough surface for the
// reference's sbpl_collision_checking sources to compile where they lie (see oracle/Makefile, target ref).
#pragma once
#include <Eigen/Dense>
#include <geometry_msgs/Point.h>
namespace tf {
// eigen_conversions: Translation * Quaternion(w, x, y, z)
inline void poseMsgToEigen(const geometry_msgs::Pose& m, Eigen::Affine3d& e)
{
    e = Eigen::Translation3d(m.position.x, m.position.y, m.position.z) *
        Eigen::Quaterniond(m.orientation.w, m.orientation.x, m.orientation.y, m.orientation.z);
}
inline void poseEigenToMsg(const Eigen::Affine3d& e, geometry_msgs::Pose& m)
{
    m.position.x = e.translation().x(); m.position.y = e.translation().y(); m.position.z = e.translation().z();
    Eigen::Quaterniond q(e.rotation());
    m.orientation.x = q.x(); m.orientation.y = q.y(); m.orientation.z = q.z(); m.orientation.w = q.w();
}
} // namespace tf
