// ORACLE (test infrastructure only -- never linked into the product path).  Stand-in for an absent third-party header,
// just enough surface for the reference's sbpl_collision_checking sources to COMPILE where they lie (make -C oracle ref ->
// oracle/_ref/libref_collision.so).  No behaviour of the hot path lives here unless the header says so.
#ifndef STUB_ROS_ROS_H
#define STUB_ROS_ROS_H
#include <ros/console.h>
#include <ros/time.h>
#include <map>
#include <string>
namespace XmlRpc {
// only named in the signatures of CollisionModelConfig::Load (collision_model_config.cpp is NOT compiled)
class XmlRpcValue { };
} // namespace XmlRpc
namespace ros {
class NodeHandle { };
} // namespace ros
#endif
