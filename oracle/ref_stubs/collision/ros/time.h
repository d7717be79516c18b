// ORACLE (test infrastructure only -- never linked into the product path).  Stand-in for an absent third-party header,
// just enough surface for the reference's sbpl_collision_checking sources to COMPILE where they lie (make -C oracle ref ->
// oracle/_ref/libref_collision.so).  No behaviour of the hot path lives here unless the header says so.
#ifndef STUB_ROS_TIME_H
#define STUB_ROS_TIME_H
namespace ros {
struct Time { double t = 0.0; static Time now() { return Time(); } Time() { } explicit Time(double s) : t(s) { } double toSec() const { return t; } };
struct Duration { double d = 0.0; Duration() { } explicit Duration(double s) : d(s) { } double toSec() const { return d; } };
inline Duration operator-(const Time& a, const Time& b) { return Duration(a.t - b.t); }
} // namespace ros
#endif
