// ORACLE (test infrastructure only -- never linked into the product path).  Stand-in for an absent third-party header,
// just enough surface for the reference's sbpl_collision_checking sources to COMPILE where they lie (make -C oracle ref ->
// oracle/_ref/libref_collision.so).  No behaviour of the hot path lives here unless the header says so.
#ifndef STUB_ROS_CONSOLE_H
#define STUB_ROS_CONSOLE_H
#include <sstream>
#include <string>
#define STUB_ROS_NOP(...) do { } while (0)
#define STUB_ROS_STREAM_NOP(args) do { if (false) { std::stringstream stub_ss__; stub_ss__ << args; } } while (0)
#define ROS_DEBUG(...) STUB_ROS_NOP()
#define ROS_INFO(...) STUB_ROS_NOP()
#define ROS_WARN(...) STUB_ROS_NOP()
#define ROS_ERROR(...) STUB_ROS_NOP()
#define ROS_FATAL(...) STUB_ROS_NOP()
#define ROS_DEBUG_NAMED(...) STUB_ROS_NOP()
#define ROS_INFO_NAMED(...) STUB_ROS_NOP()
#define ROS_WARN_NAMED(...) STUB_ROS_NOP()
#define ROS_ERROR_NAMED(...) STUB_ROS_NOP()
#define ROS_FATAL_NAMED(...) STUB_ROS_NOP()
#define ROS_DEBUG_ONCE(...) STUB_ROS_NOP()
#define ROS_WARN_ONCE(...) STUB_ROS_NOP()
#define ROS_WARN_ONCE_NAMED(...) STUB_ROS_NOP()
#define ROS_ERROR_ONCE(...) STUB_ROS_NOP()
#define ROS_DEBUG_THROTTLE(...) STUB_ROS_NOP()
#define ROS_DEBUG_COND(...) STUB_ROS_NOP()
#define ROS_DEBUG_COND_NAMED(...) STUB_ROS_NOP()
#define ROS_DEBUG_STREAM(args) STUB_ROS_STREAM_NOP(args)
#define ROS_INFO_STREAM(args) STUB_ROS_STREAM_NOP(args)
#define ROS_WARN_STREAM(args) STUB_ROS_STREAM_NOP(args)
#define ROS_ERROR_STREAM(args) STUB_ROS_STREAM_NOP(args)
#define ROS_DEBUG_STREAM_NAMED(name, args) STUB_ROS_STREAM_NOP(args)
#define ROS_INFO_STREAM_NAMED(name, args) STUB_ROS_STREAM_NOP(args)
#define ROS_WARN_STREAM_NAMED(name, args) STUB_ROS_STREAM_NOP(args)
#define ROS_ERROR_STREAM_NAMED(name, args) STUB_ROS_STREAM_NOP(args)
#define ROS_DEBUG_STREAM_COND_NAMED(c, name, args) STUB_ROS_STREAM_NOP(args)
#define ROS_ASSERT(x) STUB_ROS_NOP()
#define ROS_ASSERT_MSG(...) STUB_ROS_NOP()
#endif
