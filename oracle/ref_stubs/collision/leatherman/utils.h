// ORACLE (test infrastructure only).  Stand-in for an absent third-party header: just enough surface for the
// reference's sbpl_collision_checking sources to compile where they lie (see oracle/Makefile, target ref).
#pragma once
namespace leatherman {
inline void HSVtoRGB(double* r, double* g, double* b, double, double, double) { *r = *g = *b = 0.0; }   // marker colours only
} // namespace leatherman
#include <kdl/frames.hpp>
#include <string>
namespace leatherman {
// leatherman/utils.cpp: the segment whose joint has the given name; the index of the chain segment with that joint
inline bool getSegmentOfJoint(const KDL::Tree& tree, const std::string& joint, std::string& segment)
{
    for (const auto& e : tree.getSegments()) {
        if (e.second.segment.getJoint().getName() == joint) { segment = e.second.segment.getName(); return true; }
    }
    return false;
}
inline void printKDLChain(const KDL::Chain&, const std::string&) { }
inline bool getJointIndex(const KDL::Chain& c, const std::string& name, int& index)
{
    for (unsigned int j = 0; j < c.getNrOfSegments(); ++j) {
        if (c.getSegment(j).getJoint().getName() == name) { index = (int)j; return true; }
    }
    return false;
}
} // namespace leatherman
