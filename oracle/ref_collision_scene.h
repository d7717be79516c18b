// ORACLE (test infrastructure only).  The scene record shared by ref_collision_shim.cpp and ref_planner_shim.cpp: the
// REFERENCE's own objects (compiled from /root/reference, see ref_collision_shim.cpp) for one robot + grid.
#ifndef ORACLE_REF_COLLISION_SCENE_H
#define ORACLE_REF_COLLISION_SCENE_H

#include <memory>

#include <smpl/occupancy_grid.h>
#include <sbpl_collision_checking/collision_space.h>

#include "robot_desc.h"

namespace sbpl { class EuclidDistanceMap; }   // euclid_distance_map.h has no include guard: only the collision shim includes it

struct refcc_scene
{
    std::string robot_path;
    oracle::RobotDesc desc;
    urdf::ModelInterface urdf;
    sbpl::collision::CollisionModelConfig config;
    std::shared_ptr<sbpl::EuclidDistanceMap> df;
    std::unique_ptr<sbpl::OccupancyGrid> grid;
    std::unique_ptr<sbpl::collision::CollisionSpace> cc;
    std::vector<std::string> planning_joints;
    int dof = 0;
};

#endif
