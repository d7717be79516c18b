// ORACLE (test infrastructure only).  Stand-in for an absent third-party header: just enough surface for the
// reference's sbpl_collision_checking sources to compile where they lie (see oracle/Makefile, target ref).
#pragma once
#include <geometry_msgs/Point.h>
#include <octomap/octomap.h>
#include <std_msgs/Header.h>
#include <string>
#include <vector>
namespace octomap_msgs {
struct Octomap { std_msgs::Header header; bool binary = false; std::string id; double resolution = 0; std::vector<int8_t> data; };
struct OctomapWithPose { std_msgs::Header header; geometry_msgs::Pose origin; Octomap octomap; };
} // namespace octomap_msgs
