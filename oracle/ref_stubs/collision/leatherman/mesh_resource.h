// ORACLE (test infrastructure only).  Stand-in for an absent third-party header: just enough surface for the
// reference's sbpl_collision_checking sources to compile where they lie (see oracle/Makefile, target ref).
#pragma once
#include <Eigen/Dense>
#include <string>
#include <vector>
namespace leatherman {
// mesh files are not part of the pinned configurations (link geometry is given as boxes)
inline bool getMeshComponentsFromResource(const std::string&, const Eigen::Vector3d&, std::vector<int>&, std::vector<Eigen::Vector3d>&) { return false; }
} // namespace leatherman
