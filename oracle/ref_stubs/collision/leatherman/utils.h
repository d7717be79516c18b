// ORACLE (test infrastructure only).  Stand-in for an absent third-party header: just enough surface for the
// reference's sbpl_collision_checking sources to compile where they lie (see oracle/Makefile, target ref).
#pragma once
namespace leatherman {
inline void HSVtoRGB(double* r, double* g, double* b, double, double, double) { *r = *g = *b = 0.0; }   // marker colours only
} // namespace leatherman
