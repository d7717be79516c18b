// ORACLE (test infrastructure only -- never linked into the product path).
//
// C shim around the REFERENCE's own ARA* (smpl/src/search/arastar.cpp, compiled where it lies under
// /root/reference by `make -C oracle ref` against the interface stubs in oracle/ref_stubs/sbpl; output
// oracle/_ref/libref_arastar.so).  It runs the reference search on an explicit graph (CSR successor lists
// with integer edge costs, a heuristic value per state), configured the way our planners use it: one
// weighted-A* iteration at the initial epsilon, bounded by an expansion count.  tests/test_oracle_arastar.py
// compares oracle/arastar.h with it: same path, cost and expansion count, ties included.
#include <cstdarg>
#include <sstream>
#include <string>
#include <vector>

#include <smpl/console/console.h>
#include <smpl/search/arastar.h>

// silent definitions of the std console backend declared in smpl/console/detail/console_std.h
// (the reference's console.cpp needs Boost, which is not installed)
namespace sbpl {
namespace console {
bool g_initialized = true;
void initialize() { g_initialized = true; }
void InitializeLogLocation(LogLocation* loc, const std::string&, Level level)
{
    loc->logger = nullptr;
    loc->next = nullptr;
    loc->level = level;
    loc->enabled = false;
    loc->initialized = true;
}
void print(Level, const char*, int, const char*, ...) { }
void print(Level, const char*, int, const std::stringstream&) { }
} // namespace console
} // namespace sbpl

namespace {

struct CsrGraph : public DiscreteSpaceInformation
{
    int n;
    const int* off;
    const int* dst;
    const int* cost;
    void GetSuccs(int s, std::vector<int>* succs, std::vector<int>* costs) override
    {
        succs->clear();
        costs->clear();
        for (int k = off[s]; k < off[s + 1]; ++k) {
            succs->push_back(dst[k]);
            costs->push_back(cost[k]);
        }
    }
};

struct ArrayHeuristic : public Heuristic
{
    const int* h;
    explicit ArrayHeuristic(DiscreteSpaceInformation* g) : Heuristic(g), h(nullptr) { }
    int GetGoalHeuristic(int id) override { return h[id]; }
    int GetStartHeuristic(int) override { return 0; }
    int GetFromToHeuristic(int, int) override { return 0; }
};

} // namespace

extern "C" {

/// out[0] = replan's return value (1 = solution), out[1] = cost, out[2] = expansions, out[3] = path length
int ref_arastar_search(int n, const int* off, const int* dst, const int* cost, const int* h,
                       int start, int goal, double eps, int max_expansions,
                       int* path, int max_path, int* out)
{
    CsrGraph g;
    g.n = n; g.off = off; g.dst = dst; g.cost = cost;
    ArrayHeuristic heur(&g);
    heur.h = h;
    sbpl::ARAStar search(&g, &heur);
    search.set_initialsolution_eps(eps);
    search.set_start(start);
    search.set_goal(goal);
    sbpl::ARAStar::TimeParameters tp;
    tp.bounded = true;
    tp.improve = false;
    tp.type = sbpl::ARAStar::TimeParameters::EXPANSIONS;
    tp.max_expansions_init = max_expansions;
    tp.max_expansions = max_expansions;
    tp.max_allowed_time_init = sbpl::clock::duration::zero();
    tp.max_allowed_time = sbpl::clock::duration::zero();
    std::vector<int> solution;
    int solcost = 0;
    const int ret = search.replan(tp, &solution, &solcost);
    out[0] = ret;
    out[1] = solcost;
    out[2] = search.get_n_expands();
    out[3] = (int)solution.size();
    for (int i = 0; i < (int)solution.size() && i < max_path; ++i) {
        path[i] = solution[i];
    }
    return 0;
}

} // extern "C"
