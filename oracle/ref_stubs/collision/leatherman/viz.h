// ORACLE (test infrastructure only).  Stand-in for an absent third-party header: just enough surface for the
// reference's sbpl_collision_checking sources to compile where they lie (see oracle/Makefile, target ref).
#pragma once
#include <leatherman/utils.h>
#include <string>
#include <vector>
#include <visualization_msgs/MarkerArray.h>
namespace viz {
inline visualization_msgs::MarkerArray getSpheresMarkerArray(const std::vector<std::vector<double>>&, const std::vector<double>&, int, const std::string&, const std::string&, int) { return visualization_msgs::MarkerArray(); }
} // namespace viz
