// ORACLE (test infrastructure only).  Stand-in for an absent third-party header: just enough surface for the
// reference's sbpl_collision_checking sources to compile where they lie (see oracle/Makefile, target ref).
#pragma once
#include <geometry_msgs/Point.h>
#include <std_msgs/Header.h>
#include <string>
#include <vector>
namespace visualization_msgs {
struct Marker
{
    enum { ARROW = 0, CUBE = 1, SPHERE = 2, CYLINDER = 3, LINE_STRIP = 4, LINE_LIST = 5, CUBE_LIST = 6, SPHERE_LIST = 7,
           POINTS = 8, TEXT_VIEW_FACING = 9, MESH_RESOURCE = 10, TRIANGLE_LIST = 11 };
    enum { ADD = 0, MODIFY = 0, DELETE = 2, DELETEALL = 3 };
    std_msgs::Header header;
    std::string ns;
    int id = 0, type = 0, action = 0;
    geometry_msgs::Pose pose;
    geometry_msgs::Vector3 scale;
    std_msgs::ColorRGBA color;
    ros::Duration lifetime;
    bool frame_locked = false;
    std::vector<geometry_msgs::Point> points;
    std::vector<std_msgs::ColorRGBA> colors;
    std::string text, mesh_resource;
    bool mesh_use_embedded_materials = false;
};
struct MarkerArray { std::vector<Marker> markers; };
} // namespace visualization_msgs
