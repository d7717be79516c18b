// ORACLE (test infrastructure only).  C shim around the REFERENCE's own shortcutting templates
// (smpl/include/smpl/geometry/shortcut.h + detail/shortcut.hpp, std-only headers, compiled where they lie by
// `make -C oracle ref` into oracle/_ref/libref_shortcut.so), so that oracle/shortcut.h can be checked against them on index paths with arbitrary validity / cost tables.
// No reference source is copied: this file only instantiates the templates.
#include <cstdint>
#include <iterator>
#include <vector>

#include <smpl/geometry/shortcut.h>

namespace {

typedef int (*valid_fn)(int from, int to, void* user);
typedef double (*cost_fn)(int from, int to, void* user);

// the shape of post_processing.cpp's JointPositionShortcutPathGenerator: {start, finish} when the motion is valid
struct TableGenerator
{
    valid_fn valid;
    cost_fn cost;
    void* user;

    template <typename OutputIt>
    bool operator()(const int& start, const int& finish, OutputIt ofirst, double& c) const
    {
        if (!valid(start, finish, user)) {
            return false;
        }
        *ofirst++ = start;
        *ofirst++ = finish;
        c = cost(start, finish, user);
        return true;
    }
};

} // namespace

extern "C" {

/// algo 0 = shortcut::ShortcutPath, 1 = shortcut::DivideAndConquerShortcutPath.  costs[n - 1].
/// Returns the number of output points (indices into the input path) or -1.
int ref_shortcut_path(int n, const double* costs, valid_fn valid, cost_fn cost, void* user, int algo,
                      int granularity, int32_t* out, int max_out)
{
    std::vector<int> points(n);
    for (int i = 0; i < n; ++i) points[i] = i;
    const std::vector<int>& cpoints = points;
    std::vector<double> cvec(costs, costs + (n > 0 ? n - 1 : 0));
    const std::vector<double>& ccosts = cvec;
    TableGenerator gens[1] = { TableGenerator{ valid, cost, user } };
    std::vector<int> result;
    bool ok;
    if (algo == 0) {
        ok = sbpl::shortcut::ShortcutPath(cpoints.begin(), cpoints.end(), ccosts.begin(), ccosts.end(),
                                          gens, gens + 1, std::back_inserter(result), 1, (size_t)granularity);
    } else {
        ok = sbpl::shortcut::DivideAndConquerShortcutPath(cpoints.begin(), cpoints.end(), ccosts.begin(), ccosts.end(),
                                                          gens, gens + 1, std::back_inserter(result));
    }
    if (!ok || (int)result.size() > max_out) {
        return -1;
    }
    for (size_t i = 0; i < result.size(); ++i) out[i] = result[i];
    return (int)result.size();
}

} // extern "C"
