// ORACLE (test infrastructure only).  Stand-in for an absent third-party header: just enough surface for the
// reference's sbpl_collision_checking sources to compile where they lie (see oracle/Makefile, target ref).
#pragma once
#include <std_msgs/Header.h>
#include <string>
#include <vector>
namespace sensor_msgs {
struct JointState { std_msgs::Header header; std::vector<std::string> name; std::vector<double> position, velocity, effort; };
struct MultiDOFJointState { std_msgs::Header header; std::vector<std::string> joint_names; };
} // namespace sensor_msgs
