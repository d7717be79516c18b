// ORACLE (test infrastructure only).  Stand-in for an absent third-party header: just enough surface for the
// reference's sbpl_collision_checking sources to compile where they lie (see oracle/Makefile, target ref).
#pragma once
// geometric_shapes: plain shape records (the reference only reads their public fields).
#include <memory>
#include <vector>
namespace octomap { class OcTree; }
namespace shapes {
enum ShapeType { UNKNOWN_SHAPE, SPHERE, CYLINDER, CONE, BOX, PLANE, MESH, OCTREE };
struct Shape { ShapeType type; explicit Shape(ShapeType t = UNKNOWN_SHAPE) : type(t) { } virtual ~Shape() { } };
struct Sphere : Shape { double radius; explicit Sphere(double r = 0.0) : Shape(SPHERE), radius(r) { } };
struct Cylinder : Shape { double length, radius; Cylinder(double r = 0.0, double l = 0.0) : Shape(CYLINDER), length(l), radius(r) { } };
struct Cone : Shape { double length, radius; Cone(double r = 0.0, double l = 0.0) : Shape(CONE), length(l), radius(r) { } };
struct Box : Shape { double size[3]; Box(double x = 0.0, double y = 0.0, double z = 0.0) : Shape(BOX) { size[0] = x; size[1] = y; size[2] = z; } };
struct Plane : Shape { double a, b, c, d; Plane(double a_ = 0, double b_ = 0, double c_ = 0, double d_ = 0) : Shape(PLANE), a(a_), b(b_), c(c_), d(d_) { } };
struct Mesh : Shape
{
    unsigned int vertex_count = 0; double* vertices = nullptr; unsigned int triangle_count = 0; unsigned int* triangles = nullptr;
    std::vector<double> v_; std::vector<unsigned int> t_;
    Mesh() : Shape(MESH) { }
    Mesh(unsigned int nv, unsigned int nt) : Shape(MESH), vertex_count(nv), triangle_count(nt), v_(3 * nv), t_(3 * nt) { vertices = v_.data(); triangles = t_.data(); }
};
struct OcTree : Shape { std::shared_ptr<const octomap::OcTree> octree; OcTree() : Shape(OCTREE) { } };
typedef std::shared_ptr<Shape> ShapePtr;
typedef std::shared_ptr<const Shape> ShapeConstPtr;
} // namespace shapes
