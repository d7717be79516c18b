// ORACLE (test infrastructure only -- never linked into the product path).
//
// Restatement of the reference's in-tree lazy weighted A*:
//   sbpl::LazyARAStar / Replan   smpl/src/search/lazy_arastar.cpp:208-269 (main loop, termination test)
//   ExpandState                  :73-138    EvaluateState            :140-184
//   IsPredDominated              :64-71     ComputeFVal / StateCompare :198-204, 271-277
//   GetState / ReinitState       :14-62     ReconstructPath          :186-196
//   intrusive_heap               smpl/include/smpl/detail/intrusive_heap.hpp (push, pop, update = erase + push)
// g, h and f are `int` here (lazy_arastar.h:74-80), unlike ARAStar's unsigned: a negative heuristic sorts FIRST.
// One quirk is restated as it behaves: EvaluateState erases the only candidate of a state whose edge turned out
// invalid and then reads `*min_element(begin, end)` of the empty vector (:170-173) -- the erased element's storage,
// i.e. the values the state already has -- and does not push the state; here: the fields stay, nothing is pushed.
// The search has no expansion bound; callers bound it in the successor function (see ManipLatticePlanner::planLazy and
// oracle/ref_planner_plugins.h:LatticeLazySuccFun, which do the same thing on either side of the comparison).
// PINNED: tests/test_oracle_planner_reference.py runs this class under the oracle's lattice and the reference's own
// lazy_arastar.cpp under the reference's ManipLattice (oracle/_ref/libref_collision.so) on the same queries.
#ifndef ORACLE_LAZY_ARASTAR_H
#define ORACLE_LAZY_ARASTAR_H

#include <algorithm>
#include <functional>
#include <vector>

namespace oracle {

class LazyAraStar
{
public:
    static const int INFINITE = 1000000000;   // g_infinite, lazy_arastar.cpp:10

    typedef std::function<void(int, std::vector<int>&, std::vector<int>&, std::vector<bool>&)> LazySuccsFn;
    typedef std::function<int(int, int)> TrueCostFn;
    typedef std::function<int(int)> HeurFn;

    struct Result
    {
        bool found;
        int cost;
        std::vector<int> path;
        Result() : found(false), cost(0) { }
    };

    LazyAraStar(LazySuccsFn succs, TrueCostFn true_cost, HeurFn heur) :
        m_succs_fn(succs), m_cost_fn(true_cost), m_heur_fn(heur), m_eps(1.0), m_goal(-1) { }

    Result search(int start_id, int goal_id, double eps)
    {
        Result res;
        m_states.clear();
        m_open.assign(1, -1);
        m_eps = eps;
        m_goal = goal_id;
        touch(start_id);
        touch(goal_id);
        m_states[start_id].g = 0;
        m_states[start_id].true_cost = true;
        heapPush(start_id);

        while (m_open.size() > 1) {
            const int min_id = m_open[1];
            heapPop();
            const int fs = fval(min_id);
            if (m_states[goal_id].true_cost && fval(goal_id) <= fs) {
                m_states[goal_id].ebp = m_states[goal_id].bp;
                m_states[goal_id].eg = m_states[goal_id].g;
                for (int s = goal_id; s >= 0; s = m_states[s].ebp) {
                    res.path.push_back(s);
                }
                std::reverse(res.path.begin(), res.path.end());
                res.cost = m_states[goal_id].g;
                res.found = true;
                return res;
            }
            if (m_states[min_id].closed) {
                continue;   // a state may come up for expansion / evaluation twice
            }
            if (m_states[min_id].true_cost) {
                expand(min_id);
            } else {
                evaluate(min_id);
            }
        }
        return res;
    }

private:
    struct Cand
    {
        int pred;
        int g;
        bool true_cost;
    };

    struct State
    {
        std::vector<Cand> cands;
        int bp, ebp;
        int h, g, eg;
        bool true_cost, closed, touched;
        int heap_index;
        State() : bp(-1), ebp(-1), h(INFINITE), g(INFINITE), eg(INFINITE), true_cost(false), closed(false), touched(false), heap_index(0) { }
    };

    LazySuccsFn m_succs_fn;
    TrueCostFn m_cost_fn;
    HeurFn m_heur_fn;
    double m_eps;
    int m_goal;
    std::vector<State> m_states;
    std::vector<int> m_open;   // 1-based binary heap of state ids

    // GetState + ReinitState: the heuristic is asked the first time a state is touched in a query
    void touch(int id)
    {
        if ((int)m_states.size() <= id) {
            m_states.resize(id + 1);
        }
        State& s = m_states[id];
        if (!s.touched) {
            s.touched = true;
            s.h = (id == m_goal) ? 0 : m_heur_fn(id);
        }
    }

    int fval(int id) const { return m_states[id].g + (int)(m_eps * (double)m_states[id].h); }

    bool dominated(const Cand& c, const State& s) const
    {
        for (const Cand& o : s.cands) {
            if (o.pred != c.pred && o.true_cost && o.g <= c.g) {
                return true;
            }
        }
        return false;
    }

    static size_t best(const std::vector<Cand>& cands)   // std::min_element by g: the first of the smallest
    {
        size_t b = 0;
        for (size_t k = 1; k < cands.size(); ++k) {
            if (cands[k].g < cands[b].g) {
                b = k;
            }
        }
        return b;
    }

    void expand(int id)
    {
        m_states[id].closed = true;
        m_states[id].ebp = m_states[id].bp;
        m_states[id].eg = m_states[id].g;
        std::vector<int> succs, costs;
        std::vector<bool> trues;
        m_succs_fn(id, succs, costs, trues);
        for (size_t k = 0; k < succs.size(); ++k) {
            touch(succs[k]);
            State& ss = m_states[succs[k]];
            if (ss.closed) {
                continue;
            }
            Cand c;
            c.pred = id;
            c.g = m_states[id].g + costs[k];
            c.true_cost = trues[k];
            if (dominated(c, ss)) {
                continue;
            }
            ss.cands.push_back(c);
            const Cand& b = ss.cands[best(ss.cands)];
            ss.bp = b.pred;
            ss.g = b.g;
            ss.true_cost = b.true_cost;
            if (ss.heap_index == 0) {
                heapPush(succs[k]);
            } else {
                heapErase(succs[k]);   // intrusive_heap::update = erase + push
                heapPush(succs[k]);
            }
        }
    }

    void evaluate(int id)
    {
        State& s = m_states[id];
        size_t b = best(s.cands);
        const int cost = m_cost_fn(s.bp, id);
        if (cost < 0) {
            s.cands.erase(s.cands.begin() + b);
        } else {
            s.cands[b].true_cost = true;
            s.cands[b].g = m_states[s.cands[b].pred].g + cost;
            if (dominated(s.cands[b], s)) {
                s.cands.erase(s.cands.begin() + b);
            }
        }
        if (s.cands.empty()) {
            return;   // the reference reads the erased candidate here: the state's fields keep their values; not pushed
        }
        b = best(s.cands);
        s.bp = s.cands[b].pred;
        s.g = s.cands[b].g;
        s.true_cost = s.cands[b].true_cost;
        heapPush(id);
    }

    bool heapLess(int a, int b) const { return fval(a) < fval(b); }

    void percolateUp(size_t pivot)
    {
        const int tmp = m_open[pivot];
        while (pivot != 1) {
            const size_t p = pivot >> 1;
            if (heapLess(m_open[p], tmp)) {
                break;
            }
            m_open[pivot] = m_open[p];
            m_states[m_open[pivot]].heap_index = (int)pivot;
            pivot = p;
        }
        m_open[pivot] = tmp;
        m_states[tmp].heap_index = (int)pivot;
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
                m_states[m_open[pivot]].heap_index = (int)pivot;
                pivot = s;
            } else {
                break;
            }
            left = pivot << 1;
            right = left + 1;
        }
        m_open[pivot] = tmp;
        m_states[tmp].heap_index = (int)pivot;
    }

    void heapPush(int id)
    {
        m_states[id].heap_index = (int)m_open.size();
        m_open.push_back(id);
        percolateUp(m_open.size() - 1);
    }

    void heapPop()
    {
        m_states[m_open[1]].heap_index = 0;
        m_open[1] = m_open.back();
        m_open.pop_back();
        percolateDown(1);
    }

    void heapErase(int id)
    {
        const size_t pos = (size_t)m_states[id].heap_index;
        m_open[pos] = m_open.back();
        m_states[m_open[pos]].heap_index = (int)pos;
        m_states[id].heap_index = 0;
        m_open.pop_back();
        percolateDown(pos);
    }
};

} // namespace oracle

#endif
