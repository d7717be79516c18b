// ORACLE (test infrastructure only).  Stand-in for an absent third-party header: just enough surface for the
// reference's sbpl_collision_checking sources to compile where they lie (see oracle/Makefile, target ref).
#pragma once
// only VoxelizeOcTree (voxel_operations.cpp:405-430) names these; octrees are outside the pinned configurations
#include <memory>
namespace octomap {
struct OcTreeNode { };
class OcTree
{
public:
    struct leaf_iterator
    {
        bool operator!=(const leaf_iterator&) const { return false; }
        leaf_iterator& operator++() { return *this; }
        double getX() const { return 0; } double getY() const { return 0; } double getZ() const { return 0; } double getSize() const { return 0; }
        const OcTreeNode& operator*() const { static OcTreeNode n; return n; }
        const OcTreeNode* operator->() const { static OcTreeNode n; return &n; }
    };
    leaf_iterator begin_leafs() const { return leaf_iterator(); }
    leaf_iterator end_leafs() const { return leaf_iterator(); }
    bool isNodeOccupied(const OcTreeNode&) const { return false; }
    bool isNodeOccupied(const OcTreeNode*) const { return false; }
};
} // namespace octomap
