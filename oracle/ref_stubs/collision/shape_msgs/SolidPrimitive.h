// ORACLE (test infrastructure only).  Stand-in for an absent third-party header: just enough surface for the
// reference's sbpl_collision_checking sources to compile where they lie (see oracle/Makefile, target ref).
#pragma once
#include <geometry_msgs/Point.h>
#include <cstdint>
#include <vector>
namespace shape_msgs {
struct SolidPrimitive
{
    enum { BOX = 1, SPHERE = 2, CYLINDER = 3, CONE = 4 };
    enum { BOX_X = 0, BOX_Y = 1, BOX_Z = 2, SPHERE_RADIUS = 0, CYLINDER_HEIGHT = 0, CYLINDER_RADIUS = 1, CONE_HEIGHT = 0, CONE_RADIUS = 1 };
    uint8_t type = 0;
    std::vector<double> dimensions;
};
struct MeshTriangle { uint32_t vertex_indices[3] = { 0, 0, 0 }; };
struct Mesh { std::vector<MeshTriangle> triangles; std::vector<geometry_msgs::Point> vertices; };
struct Plane { double coef[4] = { 0, 0, 0, 0 }; };
} // namespace shape_msgs
