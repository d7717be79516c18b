// ORACLE (test infrastructure only -- never linked into the product path).  Stand-in for an absent third-party header,
// just enough surface for the reference's sbpl_collision_checking sources to COMPILE where they lie (make -C oracle ref ->
// oracle/_ref/libref_collision.so).  No behaviour of the hot path lives here unless the header says so.
#include <boost/shared_ptr.hpp>
