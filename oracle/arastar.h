// ORACLE (test infrastructure only -- never linked into the product path).
//
// Restatement of the reference's ARA* as our planners use it (one weighted-A* iteration at the initial
// epsilon, bounded by an expansion count):
//   ARAStar::replan          smpl/src/search/arastar.cpp:107-215
//   ARAStar::improvePath     :486-527        ARAStar::expand            :531-568
//   ARAStar::computeKey      :579-582        getSearchState/reinit      :586-627
//   ARAStar::extractPath     :629-640        timedOut (EXPANSIONS)      :452-480
//   intrusive_heap           smpl/include/smpl/detail/intrusive_heap.hpp (push / pop / decrease,
//                            percolate_up / percolate_down, 1-based array)
// g, h, f, eg are `unsigned int` as in the reference (arastar.h:176-187): a negative heuristic (an
// unreachable BFS cell gives cost_per_cell * -1) therefore sorts LAST in OPEN.
// PINNED: tests/test_oracle_arastar.py runs this class and the reference's own arastar.cpp (compiled by
// `make -C oracle ref` into oracle/_ref/libref_arastar.so) on the same graphs.
#ifndef ORACLE_ARASTAR_H
#define ORACLE_ARASTAR_H

#include <algorithm>
#include <functional>
#include <vector>

namespace oracle {

class AraStar
{
public:
    static const int INFINITE_COST = 1000000000; // SBPL's INFINITECOST

    typedef std::function<void(int, std::vector<int>&, std::vector<int>&)> SuccsFn;
    typedef std::function<int(int)> HeurFn;

    struct Result
    {
        bool found;
        int cost;
        int expansions;
        std::vector<int> path;
        Result() : found(false), cost(0), expansions(0) { }
    };

    AraStar(SuccsFn succs, HeurFn heur) : m_succs_fn(succs), m_heur_fn(heur), m_eps(1.0), m_iteration(1), m_call_number(0) { }

    Result search(int start_id, int goal_id, double eps, int max_expansions)
    {
        Result res;
        m_search.clear();
        m_open.assign(1, -1);
        ++m_call_number;
        m_iteration = 1;
        m_eps = eps;
        searchState(std::max(start_id, goal_id));
        reinit(m_search[start_id]);
        reinit(m_search[goal_id]);
        m_search[start_id].g = 0;
        m_search[start_id].f = computeKey(m_search[start_id]);
        heapPush(start_id);

        std::vector<int> succs, costs;
        bool found = false;
        while (m_open.size() > 1) {
            const int min_id = m_open[1];
            if (m_search[min_id].f >= m_search[goal_id].f || min_id == goal_id) {
                found = true;
                break;
            }
            if (res.expansions >= max_expansions) {
                break;
            }
            heapPop();
            m_search[min_id].iteration_closed = m_iteration;
            m_search[min_id].eg = m_search[min_id].g;

            succs.clear();
            costs.clear();
            m_succs_fn(min_id, succs, costs);
            for (size_t k = 0; k < succs.size(); ++k) {
                searchState(succs[k]);
                SearchState& ss = m_search[succs[k]];
                reinit(ss);
                const int new_cost = m_search[min_id].eg + costs[k];
                if ((unsigned int)new_cost < ss.g) {   // the reference compares int with unsigned: as unsigned
                    ss.g = new_cost;
                    ss.bp = min_id;
                    if (ss.iteration_closed != m_iteration) {
                        ss.f = computeKey(ss);
                        if (ss.heap_index != 0) {
                            percolateUp((size_t)ss.heap_index);
                        } else {
                            heapPush(ss.state_id);
                        }
                    }
                }
            }
            ++res.expansions;
        }
        if (!found) {
            return res;
        }
        // Like the reference, a path is extracted whenever improvePath reports success -- also when the goal
        // was never reached and OPEN's minimum merely has f >= INFINITECOST (only states with a negative
        // heuristic left): the "solution" is then [goal] with cost INFINITECOST.  Callers treat that cost as
        // failure (ManipLatticePlanner::plan below, BatchPlanner::finish in the product).
        for (int s = goal_id; s >= 0; s = m_search[s].bp) {
            res.path.push_back(s);
        }
        std::reverse(res.path.begin(), res.path.end());
        res.cost = (int)m_search[goal_id].g;
        res.found = true;
        return res;
    }

private:
    struct SearchState
    {
        int state_id;
        unsigned int g, h, f, eg;
        int iteration_closed, call_number;
        int bp;
        int heap_index;
    };

    SuccsFn m_succs_fn;
    HeurFn m_heur_fn;
    std::vector<SearchState> m_search;
    std::vector<int> m_open; // 1-based binary heap of state ids
    double m_eps;
    int m_iteration, m_call_number;

    void searchState(int id)
    {
        if ((int)m_search.size() <= id) {
            SearchState blank;
            blank.state_id = -1;
            blank.call_number = 0;
            blank.heap_index = 0;
            blank.g = blank.h = blank.f = blank.eg = 0;
            blank.iteration_closed = 0;
            blank.bp = -1;
            const size_t old = m_search.size();
            m_search.resize(id + 1, blank);
            for (size_t k = old; k < m_search.size(); ++k) {
                m_search[k].state_id = (int)k;
            }
        }
    }

    void reinit(SearchState& s)
    {
        if (s.call_number != m_call_number) {
            s.g = INFINITE_COST;
            s.h = m_heur_fn(s.state_id);
            s.f = INFINITE_COST;
            s.eg = INFINITE_COST;
            s.iteration_closed = 0;
            s.call_number = m_call_number;
            s.bp = -1;
        }
    }

    int computeKey(const SearchState& s) const { return s.g + (unsigned int)(m_eps * s.h); }

    bool heapLess(int a, int b) const { return m_search[a].f < m_search[b].f; }

    void percolateUp(size_t pivot)
    {
        const int tmp = m_open[pivot];
        while (pivot != 1) {
            const size_t p = pivot >> 1;
            if (heapLess(m_open[p], tmp)) {
                break;
            }
            m_open[pivot] = m_open[p];
            m_search[m_open[pivot]].heap_index = (int)pivot;
            pivot = p;
        }
        m_open[pivot] = tmp;
        m_search[tmp].heap_index = (int)pivot;
    }

    void percolateDown(size_t pivot)
    {
        if (pivot >= m_open.size()) {
            return;
        }
        size_t left = pivot << 1, right = left + 1;
        const int tmp = m_open[pivot];
        while (left < m_open.size()) {
            size_t s = right;
            if (right >= m_open.size() || heapLess(m_open[left], m_open[right])) {
                s = left;
            }
            if (heapLess(m_open[s], tmp)) {
                m_open[pivot] = m_open[s];
                m_search[m_open[pivot]].heap_index = (int)pivot;
                pivot = s;
            } else {
                break;
            }
            left = pivot << 1;
            right = left + 1;
        }
        m_open[pivot] = tmp;
        m_search[tmp].heap_index = (int)pivot;
    }

    void heapPush(int id)
    {
        m_search[id].heap_index = (int)m_open.size();
        m_open.push_back(id);
        percolateUp(m_open.size() - 1);
    }

    void heapPop()
    {
        m_search[m_open[1]].heap_index = 0;
        m_open[1] = m_open.back();
        m_open.pop_back();
        percolateDown(1);
    }
};

} // namespace oracle

#endif
