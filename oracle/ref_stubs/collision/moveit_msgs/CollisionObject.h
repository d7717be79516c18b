// ORACLE (test infrastructure only).  Stand-in for an absent third-party header: just enough surface for the
// reference's sbpl_collision_checking sources to compile where they lie (see oracle/Makefile, target ref).
#pragma once
#include <geometry_msgs/Point.h>
#include <octomap_msgs/OctomapWithPose.h>
#include <sensor_msgs/JointState.h>
#include <shape_msgs/SolidPrimitive.h>
#include <std_msgs/Header.h>
#include <string>
#include <vector>
namespace moveit_msgs {
struct CollisionObject
{
    enum { ADD = 0, REMOVE = 1, APPEND = 2, MOVE = 3 };
    std_msgs::Header header;
    std::string id;
    std::vector<shape_msgs::SolidPrimitive> primitives;
    std::vector<geometry_msgs::Pose> primitive_poses;
    std::vector<shape_msgs::Mesh> meshes;
    std::vector<geometry_msgs::Pose> mesh_poses;
    std::vector<shape_msgs::Plane> planes;
    std::vector<geometry_msgs::Pose> plane_poses;
    int8_t operation = 0;
};
struct AttachedCollisionObject { std::string link_name; CollisionObject object; std::vector<std::string> touch_links; };
struct RobotState
{
    sensor_msgs::JointState joint_state;
    sensor_msgs::MultiDOFJointState multi_dof_joint_state;
    std::vector<AttachedCollisionObject> attached_collision_objects;
};
struct AllowedCollisionEntry { std::vector<uint8_t> enabled; };
struct AllowedCollisionMatrix
{
    std::vector<std::string> entry_names;
    std::vector<AllowedCollisionEntry> entry_values;
    std::vector<std::string> default_entry_names;
    std::vector<uint8_t> default_entry_values;
};
struct PlanningSceneWorld { std::vector<CollisionObject> collision_objects; octomap_msgs::OctomapWithPose octomap; };
struct PlanningScene
{
    std::string name, robot_model_name;
    RobotState robot_state;
    AllowedCollisionMatrix allowed_collision_matrix;
    PlanningSceneWorld world;
    bool is_diff = false;
};
} // namespace moveit_msgs
