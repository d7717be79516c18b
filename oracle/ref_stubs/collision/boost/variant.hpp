// ORACLE (test infrastructure only -- never linked into the product path).  Stand-in for an absent third-party header,
// just enough surface for the reference's sbpl_collision_checking sources to COMPILE where they lie (make -C oracle ref ->
// oracle/_ref/libref_collision.so).  No behaviour of the hot path lives here unless the header says so.
#ifndef STUB_BOOST_VARIANT_HPP
#define STUB_BOOST_VARIANT_HPP
// smpl/debug/marker.h stores shapes / colours in boost::variant; the markers are visualisation only, so this
// stand-in accepts any alternative and remembers nothing.
namespace boost {
template <typename... Ts>
struct variant
{
    variant() { }
    template <typename T> variant(const T&) { }
    template <typename T> variant& operator=(const T&) { return *this; }
    int which() const { return 0; }
};
template <typename R = void> struct static_visitor { typedef R result_type; };
template <typename V, typename... Ts> typename V::result_type apply_visitor(const V&, const variant<Ts...>&) { return typename V::result_type(); }
template <typename T, typename... Ts> T get(const variant<Ts...>&) { return T(); }
} // namespace boost
#endif
