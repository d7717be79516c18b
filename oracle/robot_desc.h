// ORACLE (test infrastructure only -- never linked into the product path).
//
// Robot description record: our stand-in for the URDF + CollisionModelConfig
// pair the reference loads (collision_model_config.h:50-119; urdf::Model).
// See tools/gen_robot_fixtures.py for the file format.
#ifndef ORACLE_ROBOT_DESC_H
#define ORACLE_ROBOT_DESC_H

#include <string>
#include <utility>
#include <vector>

namespace oracle {

struct JointDesc
{
    std::string name, type, parent, child;
    double xyz[3], rpy[3], axis[3];
    bool has_limits;
    double lower, upper;
    bool has_safety;
    double soft_lower, soft_upper;
};

struct SphereConfig
{
    std::string name;
    double x, y, z, radius;
    int priority;
};

struct SpheresModelConfig
{
    std::string link_name;
    std::vector<SphereConfig> spheres;
};

struct VoxelsModelConfig
{
    std::string link_name;
    double res;
    double center[3], size[3];
};

struct GroupConfig
{
    std::string name;
    std::vector<std::string> links;
    std::vector<std::string> groups;
    std::vector<std::pair<std::string, std::string>> chains; // (base, tip)
};

struct AcmEntryDesc
{
    std::string a, b;
    bool allowed;
};

struct RobotDesc
{
    std::string name, root;
    std::string world_joint_name, world_joint_type;
    std::vector<JointDesc> joints;
    std::vector<SpheresModelConfig> spheres_models;
    std::vector<VoxelsModelConfig> voxels_models;
    std::vector<GroupConfig> groups;
    std::vector<AcmEntryDesc> acm;
};

bool LoadRobotDesc(const std::string& path, RobotDesc& desc, std::string* err = nullptr);

} // namespace oracle

#endif
