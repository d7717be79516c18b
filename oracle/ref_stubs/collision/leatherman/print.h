// ORACLE (test infrastructure only).  Stand-in for an absent third-party header: just enough surface for the
// reference's sbpl_collision_checking sources to compile where they lie (see oracle/Makefile, target ref).
#pragma once
#include <sstream>
#include <string>
#include <vector>
namespace leatherman {
template <typename T> std::string vectorToString(const std::vector<T>&) { return std::string(); }
} // namespace leatherman
template <typename T> std::string to_string(const std::vector<T>& v) { return leatherman::vectorToString(v); }
namespace std {
// leatherman streams vectors; found by ADL from inside the reference's namespaces
template <typename T> ostream& operator<<(ostream& o, const vector<T>& v)
{
    o << "[ "; for (const T& e : v) o << e << ' '; o << ']'; return o;
}
} // namespace std
