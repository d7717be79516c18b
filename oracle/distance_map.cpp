// ORACLE (test infrastructure only -- never linked into the product path).
#include "distance_map.h"

#include <algorithm>
#include <cassert>
#include <cmath>

namespace oracle {

/// distance_map_common.h:55-58
static inline int dirnum(int dx, int dy, int dz, int edge = 0)
{
    return 27 * edge + 9 * (dx + 1) + 3 * (dy + 1) + (dz + 1);
}

/// distance_map_common.cpp:39-84
static void CreateNeighborUpdateList(
    std::array<std::array<int, 3>, 27>& neighbors,
    std::array<int, 776>& indices,
    std::array<std::pair<int, int>, 54>& ranges)
{
    int i = 0;
    int n = 0;
    for (int edge = 0; edge < 2; ++edge) {
    for (int sx = -1; sx <= 1; ++sx) {
    for (int sy = -1; sy <= 1; ++sy) {
    for (int sz = -1; sz <= 1; ++sz) {
        if (!edge) {
            neighbors[n++] = {{ sx, sy, sz }};
        }
        int d = dirnum(sx, sy, sz, edge);
        int nfirst = i;
        for (int tx = -1; tx <= 1; ++tx) {
        for (int ty = -1; ty <= 1; ++ty) {
        for (int tz = -1; tz <= 1; ++tz) {
            if (tx == 0 && ty == 0 && tz == 0) {
                continue;
            }
            if (edge) {
                if (!(((tx == -1 && sx == 1) || (tx == 1 && sx == -1)) ||
                      ((ty == -1 && sy == 1) || (ty == 1 && sy == -1)) ||
                      ((tz == -1 && sz == 1) || (tz == 1 && sz == -1))))
                {
                    indices[i++] = dirnum(tx, ty, tz);
                }
            } else {
                if (tx * sx + ty * sy + tz * sz >= 0) {
                    indices[i++] = dirnum(tx, ty, tz);
                }
            }
        }
        }
        }
        int nlast = i;
        ranges[d].first = nfirst;
        ranges[d].second = nlast;
    }
    }
    }
    }
    assert(i == 776);
}

/// distance_map.hpp:111-180
EuclidDistanceMap::EuclidDistanceMap(
    double origin_x, double origin_y, double origin_z,
    double size_x, double size_y, double size_z,
    double resolution, double max_dist)
:
    lookups(0),
    m_origin_x(origin_x), m_origin_y(origin_y), m_origin_z(origin_z),
    m_size_x(size_x), m_size_y(size_y), m_size_z(size_z), m_res(resolution),
    m_max_dist(max_dist),
    m_inv_res(1.0 / resolution),
    m_dmax_int((int)std::ceil(m_max_dist * m_inv_res)),
    m_dmax_sqrd_int(m_dmax_int * m_dmax_int),
    m_bucket(m_dmax_sqrd_int + 1),
    m_no_update_dir(dirnum(0, 0, 0))
{
    m_nx = (int)(size_x * m_inv_res + 0.5) + 2;
    m_ny = (int)(size_y * m_inv_res + 0.5) + 2;
    m_nz = (int)(size_z * m_inv_res + 0.5) + 2;

    m_open.resize(m_dmax_sqrd_int + 1);

    m_sqrt_table.resize(m_dmax_sqrd_int + 1, 0.0);
    for (int i = 0; i < m_dmax_sqrd_int + 1; ++i) {
        m_sqrt_table[i] = m_res * std::sqrt((double)i);
    }

    CreateNeighborUpdateList(m_neighbors, m_indices, m_neighbor_ranges);

    for (size_t i = 0; i < m_indices.size(); ++i) {
        const std::array<int, 3>& neighbor = m_neighbors[m_indices[i]];
        m_neighbor_offsets[i] = 0;
        m_neighbor_offsets[i] += neighbor[0] * m_nz * m_ny;
        m_neighbor_offsets[i] += neighbor[1] * m_nz;
        m_neighbor_offsets[i] += neighbor[2] * 1;
        if ((int)i < NON_BORDER_NEIGHBOR_LIST_SIZE) {
            m_neighbor_dirs[i] = dirnum(neighbor[0], neighbor[1], neighbor[2]);
        } else {
            m_neighbor_dirs[i] = dirnum(neighbor[0], neighbor[1], neighbor[2], 1);
        }
    }

    m_cells.resize((size_t)m_nx * m_ny * m_nz);
    // upstream semantics (fork defect 1): reset + coordinates for interior cells
    for (int x = 1; x < m_nx - 1; ++x) {
    for (int y = 1; y < m_ny - 1; ++y) {
    for (int z = 1; z < m_nz - 1; ++z) {
        Cell& c = cell(x, y, z);
        resetCell(c);
        c.x = x;
        c.y = y;
        c.z = z;
    }
    }
    }
    initBorderCells();
    propagateBorder();
}

/// distance_map.hpp:841-852
void EuclidDistanceMap::resetCell(Cell& c) const
{
    c.dist = m_dmax_sqrd_int;
    c.dist_new = m_dmax_sqrd_int;
    c.obs = nullptr;
    c.bucket = -1;
    c.dir = m_no_update_dir;
}

/// distance_map.hpp:432-447
void EuclidDistanceMap::reset()
{
    for (int x = 1; x < m_nx - 1; ++x) {
    for (int y = 1; y < m_ny - 1; ++y) {
    for (int z = 1; z < m_nz - 1; ++z) {
        resetCell(cell(x, y, z));
    }
    }
    }
    initBorderCells();
    propagateBorder();
}

/// distance_map.hpp:560-602
void EuclidDistanceMap::initBorderCells()
{
    auto init_obs_cell = [&](int x, int y, int z) {
        Cell& c = cell(x, y, z);
        c.x = x;
        c.y = y;
        c.z = z;
        c.dist = m_dmax_sqrd_int;
        c.dist_new = 0;
        c.obs = &c;
        c.bucket = -1;
        int src_dir_x = (x == 0) ? 1 : ((x == m_nx - 1) ? -1 : 0);
        int src_dir_y = (y == 0) ? 1 : ((y == m_ny - 1) ? -1 : 0);
        int src_dir_z = (z == 0) ? 1 : ((z == m_nz - 1) ? -1 : 0);
        c.dir = dirnum(src_dir_x, src_dir_y, src_dir_z, 1);
        updateVertex(&c);
    };
    for (int y = 0; y < m_ny; ++y) {
    for (int z = 0; z < m_nz; ++z) {
        init_obs_cell(0, y, z);
        init_obs_cell(m_nx - 1, y, z);
    }
    }
    for (int x = 1; x < m_nx - 1; ++x) {
    for (int z = 0; z < m_nz; ++z) {
        init_obs_cell(x, 0, z);
        init_obs_cell(x, m_ny - 1, z);
    }
    }
    for (int x = 1; x < m_nx - 1; ++x) {
    for (int y = 1; y < m_ny - 1; ++y) {
        init_obs_cell(x, y, 0);
        init_obs_cell(x, y, m_nz - 1);
    }
    }
}

/// distance_map.hpp:604-618 with the VECTOR_BUCKET_LIST_* macros (:44-72)
void EuclidDistanceMap::updateVertex(Cell* o)
{
    const int key = std::min(o->dist, o->dist_new);
    if (o->bucket >= 0) { // in heap: BUCKET_UPDATE
        m_open[o->bucket][o->pos] = m_open[o->bucket].back();
        m_open[o->bucket][o->pos]->pos = o->pos;
        m_open[o->bucket].pop_back();
        o->pos = (int)m_open[key].size();
        m_open[key].push_back(o);
        o->bucket = key;
    } else { // BUCKET_INSERT
        o->pos = (int)m_open[key].size();
        m_open[key].push_back(o);
        o->bucket = key;
    }
    if (key < m_bucket) {
        m_bucket = key;
    }
}

/// euclid_distance_map.cpp:49-56
int EuclidDistanceMap::distance(const Cell& n, const Cell& s) const
{
    int dx = n.x - s.obs->x;
    int dy = n.y - s.obs->y;
    int dz = n.z - s.obs->z;
    return dx * dx + dy * dy + dz * dz;
}

/// distance_map.hpp:626-643
void EuclidDistanceMap::lower(Cell* s)
{
    int nfirst = m_neighbor_ranges[s->dir].first;
    int nlast = m_neighbor_ranges[s->dir].second;
    for (int i = nfirst; i != nlast; ++i) {
        Cell* n = s + m_neighbor_offsets[i];
        int dp = distance(*n, *s);
        if (dp < n->dist_new) {
            n->dist_new = dp;
            n->obs = s->obs;
            n->dir = m_neighbor_dirs[i];
            updateVertex(n);
        }
    }
}

/// distance_map.hpp:683-694
void EuclidDistanceMap::raise(Cell* s)
{
    int nfirst = m_neighbor_ranges[m_no_update_dir].first;
    int nlast = m_neighbor_ranges[m_no_update_dir].second;
    for (int i = nfirst; i != nlast; ++i) {
        Cell* n = s + m_neighbor_offsets[i];
        waveout(n);
    }
    waveout(s);
}

/// distance_map.hpp:696-726
void EuclidDistanceMap::waveout(Cell* n)
{
    if (n == n->obs) {
        return;
    }
    n->dist_new = m_dmax_sqrd_int;
    Cell* obs_old = n->obs;
    n->obs = nullptr;

    int nfirst = m_neighbor_ranges[m_no_update_dir].first;
    int nlast = m_neighbor_ranges[m_no_update_dir].second;
    for (int i = nfirst; i != nlast; ++i) {
        Cell* a = n + m_neighbor_offsets[i];
        if (a->obs && a->obs->obs == a->obs) {
            int dp = distance(*n, *a);
            if (dp < n->dist_new) {
                n->dist_new = dp;
                n->obs = a->obs;
                n->dir = m_no_update_dir;
            }
        }
    }
    if (n->obs != obs_old) {
        updateVertex(n);
    }
}

/// distance_map.hpp:728-762
void EuclidDistanceMap::propagate()
{
    while (m_bucket < (int)m_open.size()) {
        while (!m_open[m_bucket].empty()) {
            Cell* s = m_open[m_bucket].back();
            m_open[m_bucket].pop_back();
            s->bucket = -1;

            if (s->dist_new < s->dist) {
                s->dist = s->dist_new;
                lower(s);
            } else {
                s->dist = m_dmax_sqrd_int;
                s->dir = m_no_update_dir;
                raise(s);
                if (s->dist != s->dist_new) {
                    updateVertex(s);
                }
            }
        }
        ++m_bucket;
    }
}

/// distance_map.hpp:764-781
void EuclidDistanceMap::lowerBounded(Cell* s)
{
    int nfirst = m_neighbor_ranges[s->dir].first;
    int nlast = m_neighbor_ranges[s->dir].second;
    for (int i = nfirst; i != nlast; ++i) {
        Cell* n = s + m_neighbor_offsets[i];
        if (n->dist_new > s->dist_new) {
            int dp = distance(*n, *s);
            if (dp < n->dist_new) {
                n->dist_new = dp;
                n->obs = s->obs;
                n->dir = m_neighbor_dirs[i];
                updateVertex(n);
            }
        }
    }
}

/// distance_map.hpp:783-812
void EuclidDistanceMap::propagateRemovals()
{
    while (!m_rem_stack.empty()) {
        Cell* s = m_rem_stack.back();
        m_rem_stack.pop_back();

        int nfirst = m_neighbor_ranges[m_no_update_dir].first;
        int nlast = m_neighbor_ranges[m_no_update_dir].second;
        for (int i = nfirst; i != nlast; ++i) {
            Cell* n = s + m_neighbor_offsets[i];
            bool valid = n->obs && n->obs->obs == n->obs;
            if (!valid) {
                if (n->dist_new != m_dmax_sqrd_int) {
                    n->dist_new = m_dmax_sqrd_int;
                    n->dist = m_dmax_sqrd_int;
                    n->obs = nullptr;
                    n->dir = m_no_update_dir;
                    m_rem_stack.push_back(n);
                }
            } else {
                updateVertex(n);
            }
        }
    }
    propagateBorder();
}

/// distance_map.hpp:814-839
void EuclidDistanceMap::propagateBorder()
{
    while (m_bucket < (int)m_open.size()) {
        while (!m_open[m_bucket].empty()) {
            Cell* s = m_open[m_bucket].back();
            m_open[m_bucket].pop_back();
            s->bucket = -1;
            s->dist = s->dist_new;
            lowerBounded(s);
        }
        ++m_bucket;
    }
}

/// distance_map.hpp:305-328
void EuclidDistanceMap::addPointsToMap(const std::vector<Vec3>& points)
{
    for (const Vec3& p : points) {
        int gx, gy, gz;
        worldToGrid(p.x, p.y, p.z, gx, gy, gz);
        if (!isCellValid(gx, gy, gz)) {
            continue;
        }
        ++gx; ++gy; ++gz;
        Cell& c = cell(gx, gy, gz);
        if (c.dist_new > 0) {
            c.dir = m_no_update_dir;
            c.dist_new = 0;
            c.obs = &c;
            updateVertex(&c);
        }
    }
    propagate();
}

void EuclidDistanceMap::addCellsToMap(const std::vector<std::array<int, 3>>& cells)
{
    for (const auto& g : cells) {
        int gx = g[0], gy = g[1], gz = g[2];
        if (!isCellValid(gx, gy, gz)) {
            continue;
        }
        ++gx; ++gy; ++gz;
        Cell& c = cell(gx, gy, gz);
        if (c.dist_new > 0) {
            c.dir = m_no_update_dir;
            c.dist_new = 0;
            c.obs = &c;
            updateVertex(&c);
        }
    }
    propagate();
}

/// distance_map.hpp:334-361
void EuclidDistanceMap::removePointsFromMap(const std::vector<Vec3>& points)
{
    for (const Vec3& p : points) {
        int gx, gy, gz;
        worldToGrid(p.x, p.y, p.z, gx, gy, gz);
        if (!isCellValid(gx, gy, gz)) {
            continue;
        }
        ++gx; ++gy; ++gz;
        Cell& c = cell(gx, gy, gz);
        if (c.obs != &c) {
            continue;
        }
        c.dist_new = m_dmax_sqrd_int;
        c.obs = nullptr;
        c.dist = m_dmax_sqrd_int;
        c.dir = m_no_update_dir;
        m_rem_stack.push_back(&c);
    }
    propagateRemovals();
}

/// distance_map.hpp:281-286
double EuclidDistanceMap::getDistance(double x, double y, double z) const
{
    ++lookups;
    int gx, gy, gz;
    worldToGrid(x, y, z, gx, gy, gz);
    return getDistance(gx, gy, gz);
}

/// distance_map.hpp:292-300
double EuclidDistanceMap::getDistance(int x, int y, int z) const
{
    if (!isCellValid(x, y, z)) {
        return 0.0;
    }
    int d2 = cell(x + 1, y + 1, z + 1).dist;
    return m_sqrt_table[d2];
}

int EuclidDistanceMap::getSquaredCellDistance(int x, int y, int z) const
{
    if (!isCellValid(x, y, z)) {
        return 0;
    }
    return cell(x + 1, y + 1, z + 1).dist;
}

/// distance_map.hpp:507-514
void EuclidDistanceMap::gridToWorld(int x, int y, int z, double& wx, double& wy, double& wz) const
{
    wx = (m_origin_x - m_res) + (x + 1) * m_res;
    wy = (m_origin_y - m_res) + (y + 1) * m_res;
    wz = (m_origin_z - m_res) + (z + 1) * m_res;
}

/// distance_map.hpp:519-527
void EuclidDistanceMap::worldToGrid(double wx, double wy, double wz, int& x, int& y, int& z) const
{
    x = (int)(m_inv_res * (wx - (m_origin_x - m_res)) + 0.5) - 1;
    y = (int)(m_inv_res * (wy - (m_origin_y - m_res)) + 0.5) - 1;
    z = (int)(m_inv_res * (wz - (m_origin_z - m_res)) + 0.5) - 1;
}

/// distance_map.hpp:531-536
bool EuclidDistanceMap::isCellValid(int x, int y, int z) const
{
    return x >= 0 && x < m_nx - 2 && y >= 0 && y < m_ny - 2 && z >= 0 && z < m_nz - 2;
}

} // namespace oracle
