// ORACLE (test infrastructure only -- never linked into the product path).
//
// CPU restatement of the path post-processing that re-checks motions through the hot path
// (SURVEY.md section 8f row 4):
//   shortcut::ShortcutPath               smpl/include/smpl/geometry/detail/shortcut.hpp:112-283
//   shortcut::DivideAndConquerShortcutPath    same file :285-438
//   distance / pv_distance               smpl/src/post_processing.cpp:48-97
//   JointPositionShortcutPathGenerator, JointPositionVelocityShortcutPathGenerator   :99-161
//   ShortcutPath(rm, cc, pin, pout, type)     :284-365  (JOINT_SPACE, JOINT_POSITION_VELOCITY_SPACE;
//                                             EUCLID_SPACE needs third-party IK and is outside the parity set)
//   CreatePositionVelocityPath           :367-401
//   InterpolatePath                      :476-540
//
// Pinned: the two shortcut templates are checked against the reference's own header compiled here
// (oracle/ref_shortcut_shim.cpp -> oracle/_ref/libref_shortcut.so, tests/test_oracle_shortcut.py).  The
// generators and cost functions around them are restated from post_processing.cpp, which needs Eigen /
// MoveIt types and cannot be compiled here.
//
// Both joint-space generators answer a (start, finish) request with the two-point path {start, finish} when
// isStateToStateValid(start, finish) holds, so every shortcut path is a subsequence of the input path: the
// functions below work on point INDICES.
#ifndef ORACLE_SHORTCUT_H
#define ORACLE_SHORTCUT_H

#include <algorithm>
#include <cmath>
#include <functional>
#include <vector>

namespace oracle {

/// a path generator: true + cost when it has a path from point `from` to point `to` (always the two end points)
typedef std::function<bool(int from, int to, double& cost)> ShortcutGenerator;

namespace shortcut_detail {

inline std::vector<int> IndexRange(int first, int last)
{
    std::vector<int> r;
    for (int i = first; i <= last; ++i) r.push_back(i);
    return r;
}

inline std::vector<double> Accumulate(const std::vector<double>& costs)
{
    std::vector<double> accum(costs.size() + 1);
    accum[0] = 0.0;
    for (size_t i = 1; i < accum.size(); ++i) {
        accum[i] = accum[i - 1] + costs[i - 1];
    }
    return accum;
}

} // namespace shortcut_detail

/// shortcut.hpp:112-283.  n points, costs[n - 1]; `window` is unused by the reference as well.
inline bool ShortcutPath(int n, const std::vector<double>& costs, const std::vector<ShortcutGenerator>& gens,
                         std::vector<int>& out, size_t granularity = 1)
{
    using namespace shortcut_detail;
    const size_t psize = (size_t)n;
    if (psize == 0) {
        return true;
    }
    if (psize != costs.size() + 1) {
        return false;
    }
    if (psize < 2) {
        out.push_back(0);
        return true;
    }
    const std::vector<double> accum = Accumulate(costs);

    // the segment of interest [seg_start, seg_end] and the best known path over it
    size_t seg_start = 0;
    size_t seg_end = std::min(psize - 1, granularity);
    std::vector<int> best;
    double best_cost = 0.0;
    auto open_segment = [&]() {
        best = IndexRange((int)seg_start, (int)seg_end);
        best_cost = accum[seg_end] - accum[seg_start];
        for (const ShortcutGenerator& g : gens) {
            double cost;
            if (g((int)seg_start, (int)seg_end, cost) && cost <= best_cost) {
                best = { (int)seg_start, (int)seg_end };
                best_cost = cost;
            }
        }
    };
    open_segment();
    out.push_back(0);

    while (seg_end != psize) {   // psize stands for the reference's "curr_end == plast"
        bool improved = false;
        const size_t look = std::min(granularity, psize - seg_end - 1);
        if (look != 0) {
            double new_cost = best_cost + (accum[seg_end + look] - accum[seg_end]);
            for (const ShortcutGenerator& g : gens) {
                double cost;
                if (g((int)seg_start, (int)(seg_end + look), cost) && cost <= new_cost) {
                    improved = true;
                    best = { (int)seg_start, (int)(seg_end + look) };
                    new_cost = cost;
                }
            }
            best_cost = new_cost;
        }
        if (improved) {
            seg_end += look;
        } else if (look == 0) {
            seg_end = psize;   // the final best path is emitted after the loop
        } else {
            out.insert(out.end(), best.begin() + 1, best.end());
            seg_start = seg_end;
            seg_end += look;
            open_segment();
        }
    }
    out.insert(out.end(), best.begin() + 1, best.end());
    return true;
}

/// shortcut.hpp:285-438
inline bool DivideAndConquerShortcutPath(int n, const std::vector<double>& costs,
                                         const std::vector<ShortcutGenerator>& gens, std::vector<int>& out)
{
    using namespace shortcut_detail;
    const size_t psize = (size_t)n;
    if (psize == 0) {
        return true;
    }
    if (psize != costs.size() + 1) {
        return false;
    }
    if (psize < 2) {
        out.push_back(0);
        return true;
    }
    const std::vector<double> accum = Accumulate(costs);
    std::function<void(int, int)> rec = [&](int first, int last) {
        if (last - first == 1) {
            out.push_back(last);
            return;
        }
        bool improved = false;
        double best_cost = accum[last] - accum[first];
        for (const ShortcutGenerator& g : gens) {
            double cost;
            if (g(first, last, cost) && cost <= best_cost) {
                improved = true;
                best_cost = cost;
            }
        }
        if (improved) {
            out.push_back(last);
            return;
        }
        const int mid = first + ((last - first) >> 1);
        rec(first, mid);
        rec(mid, last);
    };
    out.push_back(0);
    rec(0, n - 1);
    return true;
}

/// angles.h:45-99 (kept local so this header stands alone)
inline double ShortcutNormalizeAngle(double angle)
{
    if (std::fabs(angle) > 2.0 * M_PI) {
        angle = std::fmod(angle, 2.0 * M_PI);
    }
    if (angle < -M_PI) {
        angle += 2.0 * M_PI;
    }
    if (angle > M_PI) {
        angle -= 2.0 * M_PI;
    }
    return angle;
}

/// post_processing.cpp:48-67: `continuous` = !RobotModel::hasPosLimit
inline double JointDistance(const std::vector<bool>& continuous, const double* from, const double* to)
{
    double dist = 0.0;
    for (size_t v = 0; v < continuous.size(); ++v) {
        if (continuous[v]) {
            dist += std::fabs(ShortcutNormalizeAngle(to[v] - from[v]));
        } else {
            dist += std::fabs(to[v] - from[v]);
        }
    }
    return dist;
}

enum ShortcutKind { SHORTCUT_JOINT_SPACE = 0, SHORTCUT_JOINT_POSITION_VELOCITY_SPACE = 1 };

/// post_processing.cpp:284-365.  path: n x dof joint positions; valid(i, j) = isStateToStateValid(path[i], path[j]).
/// Returns the indices of the shortcut path's points.
inline std::vector<int> ShortcutJointPath(const std::vector<bool>& continuous, const double* path, int n,
                                          const std::function<bool(int, int)>& valid, int kind)
{
    const size_t dof = continuous.size();
    std::vector<int> out;
    if (n < 2) {
        for (int i = 0; i < n; ++i) out.push_back(i);
        return out;
    }
    auto pt = [&](int i) { return path + (size_t)i * dof; };
    if (kind == SHORTCUT_JOINT_SPACE) {
        std::vector<double> costs(n - 1);
        for (int i = 1; i < n; ++i) {
            costs[i - 1] = JointDistance(continuous, pt(i - 1), pt(i));
        }
        std::vector<ShortcutGenerator> gens = {
            [&](int a, int b, double& cost) {
                if (!valid(a, b)) return false;
                cost = JointDistance(continuous, pt(a), pt(b));
                return true;
            } };
        ShortcutPath(n, costs, gens, out);
        return out;
    }
    // CreatePositionVelocityPath (:367-401): velocity = sign of the motion into the point, 0 for the first
    std::vector<double> vel((size_t)n * dof, 0.0);
    for (int i = 1; i < n; ++i) {
        for (size_t v = 0; v < dof; ++v) {
            const double d = continuous[v] ? ShortcutNormalizeAngle(pt(i)[v] - pt(i - 1)[v]) : pt(i)[v] - pt(i - 1)[v];
            vel[(size_t)i * dof + v] = std::copysign(1.0, d);
        }
    }
    // pv_distance (:69-97)
    auto pv = [&](int a, int b) {
        const double dist = JointDistance(continuous, pt(a), pt(b));
        double vdist = 0.0;
        for (size_t v = 0; v < dof; ++v) {
            vdist += std::fabs(vel[(size_t)b * dof + v] - vel[(size_t)a * dof + v]);
        }
        return 1.0 * dist + 1.0 * vdist;
    };
    std::vector<double> costs(n - 1);
    for (int i = 1; i < n; ++i) {
        costs[i - 1] = pv(i - 1, i);
    }
    std::vector<ShortcutGenerator> gens = {
        [&](int a, int b, double& cost) {
            if (!valid(a, b)) return false;
            cost = pv(a, b);
            return true;
        } };
    std::vector<int> greedy, dnc;
    ShortcutPath(n, costs, gens, greedy);
    DivideAndConquerShortcutPath(n, costs, gens, dnc);
    auto total = [&](const std::vector<int>& idx) {
        // ComputePositionVelocityPathCosts + std::accumulate over the shortcut pv path (the velocities travel
        // with the points)
        double c = 0.0;
        for (size_t i = 1; i < idx.size(); ++i) {
            c = c + pv(idx[i - 1], idx[i]);
        }
        return c;
    };
    return total(dnc) < total(greedy) ? dnc : greedy;
}

} // namespace oracle

#endif
