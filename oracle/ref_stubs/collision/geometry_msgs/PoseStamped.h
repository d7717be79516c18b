// ORACLE (test infrastructure only).  Stand-in for an absent third-party header (see geometry_msgs/Point.h).
#pragma once
#include <geometry_msgs/Point.h>
#include <ros/console.h>
#include <std_msgs/Header.h>
namespace geometry_msgs {
struct PoseStamped { std_msgs::Header header; Pose pose; };
} // namespace geometry_msgs
