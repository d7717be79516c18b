// ORACLE (test infrastructure only).  Stand-in for an absent third-party header: just enough surface for the
// reference's sbpl_collision_checking sources to compile where they lie (see oracle/Makefile, target ref).
#pragma once
#include <geometric_shapes/shapes.h>
#include <shape_msgs/SolidPrimitive.h>
namespace shapes {
// only types.cpp (ConvertCollisionObjectToObject, not on the checked path) calls these
inline Shape* constructShapeFromMsg(const shape_msgs::SolidPrimitive&) { return nullptr; }
inline Shape* constructShapeFromMsg(const shape_msgs::Mesh&) { return nullptr; }
inline Shape* constructShapeFromMsg(const shape_msgs::Plane&) { return nullptr; }
template <typename M> bool constructMarkerFromShape(const Shape*, M&, bool = false) { return false; }   // visualisation only
} // namespace shapes
