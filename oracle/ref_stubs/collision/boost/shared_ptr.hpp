// ORACLE (test infrastructure only -- never linked into the product path).  Stand-in for an absent third-party header,
// just enough surface for the reference's sbpl_collision_checking sources to COMPILE where they lie (make -C oracle ref ->
// oracle/_ref/libref_collision.so).  No behaviour of the hot path lives here unless the header says so.
#ifndef STUB_BOOST_SHARED_PTR_HPP
#define STUB_BOOST_SHARED_PTR_HPP
#include <memory>
namespace boost {
using std::shared_ptr;
using std::weak_ptr;
using std::make_shared;
using std::static_pointer_cast;
using std::dynamic_pointer_cast;
using std::const_pointer_cast;
} // namespace boost
#endif
