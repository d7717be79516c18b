// ORACLE (test infrastructure only).  Stand-in for an absent third-party header (see OctomapWithPose.h).
#pragma once
#include <octomap_msgs/OctomapWithPose.h>
namespace octomap { class AbstractOcTree { public: virtual ~AbstractOcTree() { } }; }
namespace octomap_msgs {
inline octomap::AbstractOcTree* fullMsgToMap(const Octomap&) { return nullptr; }
inline octomap::AbstractOcTree* binaryMsgToMap(const Octomap&) { return nullptr; }
inline octomap::AbstractOcTree* msgToMap(const Octomap&) { return nullptr; }
} // namespace octomap_msgs
