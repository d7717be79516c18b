// ORACLE (test infrastructure only).  Stand-in for an absent third-party header: just enough surface for the
// reference's sbpl_collision_checking sources to compile where they lie (see oracle/Makefile, target ref).
#pragma once
#include <Eigen/Dense>
#include <Eigen/StdVector>
#include <geometric_shapes/shapes.h>
#include <memory>
#include <string>
#include <vector>
namespace collision_detection {
class World
{
public:
    struct Object
    {
        explicit Object(const std::string& id) : id_(id) { }
        std::string id_;
        std::vector<shapes::ShapeConstPtr> shapes_;
        std::vector<Eigen::Affine3d, Eigen::aligned_allocator<Eigen::Affine3d>> shape_poses_;
    };
    typedef std::shared_ptr<Object> ObjectPtr;
    typedef std::shared_ptr<const Object> ObjectConstPtr;
};
} // namespace collision_detection
