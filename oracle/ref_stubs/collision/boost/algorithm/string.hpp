// ORACLE (test infrastructure only).  Stand-in for an absent third-party header (named by
// manip_lattice_action_space.h; nothing on the compiled path calls it).
#pragma once
