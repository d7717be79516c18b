// ORACLE (test infrastructure only -- never linked into the product path).
#include "bfs3d.h"

namespace oracle {

/// bfs3d.cpp:40-111
BFS_3D::BFS_3D(int width, int height, int length) :
    expansions(0),
    m_dim_x(0), m_dim_y(0), m_dim_z(0), m_dim_xy(0), m_dim_xyz(0),
    m_queue_head(0), m_queue_tail(0)
{
    if (width <= 0 || height <= 0 || length <= 0) {
        return;
    }
    m_dim_x = width + 2;
    m_dim_y = height + 2;
    m_dim_z = length + 2;
    m_dim_xy = m_dim_x * m_dim_y;
    m_dim_xyz = m_dim_xy * m_dim_z;

    const int w = m_dim_x, p = m_dim_xy;
    const int offs[26] = {
        -w, 1, w, -1, -w - 1, -w + 1, w + 1, w - 1,
        p, -w + p, 1 + p, w + p, -1 + p, -w - 1 + p, -w + 1 + p, w + 1 + p, w - 1 + p,
        -p, -w - p, 1 - p, w - p, -1 - p, -w - 1 - p, -w + 1 - p, w + 1 - p, w - 1 - p,
    };
    for (int i = 0; i < 26; ++i) {
        m_neighbor_offsets[i] = offs[i];
    }

    m_distance_grid.resize(m_dim_xyz);
    m_queue.resize((size_t)width * height * length);

    for (int node = 0; node < m_dim_xyz; node++) {
        int x = node % m_dim_x;
        int y = node / m_dim_x % m_dim_y;
        int z = node / m_dim_xy;
        if (x == 0 || x == m_dim_x - 1 ||
            y == 0 || y == m_dim_y - 1 ||
            z == 0 || z == m_dim_z - 1)
        {
            m_distance_grid[node] = WALL;
        } else {
            m_distance_grid[node] = UNDISCOVERED;
        }
    }
}

bool BFS_3D::inBounds(int x, int y, int z) const
{
    return !(x < 0 || y < 0 || z < 0 ||
            x >= m_dim_x - 2 || y >= m_dim_y - 2 || z >= m_dim_z - 2);
}

int BFS_3D::getNode(int x, int y, int z) const
{
    if (!inBounds(x, y, z)) {
        return -1;
    }
    return (z + 1) * m_dim_xy + (y + 1) * m_dim_x + (x + 1);
}

/// bfs3d.cpp:132-141 (the reference indexes with getNode() == -1 when out of
/// bounds -- undefined behaviour; callers only pass in-bounds cells)
void BFS_3D::setWall(int x, int y, int z)
{
    int node = getNode(x, y, z);
    if (node < 0) {
        return;
    }
    m_distance_grid[node] = WALL;
}

bool BFS_3D::isWall(int x, int y, int z) const
{
    int node = getNode(x, y, z);
    return node >= 0 && m_distance_grid[node] == WALL;
}

/// bfs3d.cpp:156-201
int BFS_3D::run(int x, int y, int z)
{
    for (int i = 0; i < m_dim_xyz; i++) {
        if (m_distance_grid[i] != WALL) {
            m_distance_grid[i] = UNDISCOVERED;
        }
    }
    int origin = getNode(x, y, z);
    if (origin == -1) {
        return 0;
    }
    m_queue_head = 0;
    m_queue_tail = 1;
    m_queue[0] = origin;
    m_distance_grid[origin] = 0;
    search();
    return 1;
}

/// bfs3d.h:157-211.  `xyz` holds count*3 ints.  Faithful to the reference's
/// iterator walk: a triple is committed only when a further element follows
/// it, so the LAST triple of the range is never seeded.
int BFS_3D::run(const int* cells, int count)
{
    int numGoals = 0;
    for (int i = 0; i < m_dim_xyz; i++) {
        if (m_distance_grid[i] != WALL) {
            m_distance_grid[i] = UNDISCOVERED;
        }
    }
    m_queue_head = 0;
    int xyz[3];
    int ind = 0;
    int start_count = 0;
    const int n = count * 3;
    for (int it = 0; it != n; ++it) {
        if (ind == 3) {
            const int origin = getNode(xyz[0], xyz[1], xyz[2]);
            if (origin != -1) {
                numGoals++;
                m_queue[start_count++] = origin;
                m_distance_grid[origin] = 0;
            }
            ind = 0;
            it--;
        } else {
            xyz[ind++] = cells[it];
        }
    }
    m_queue_tail = start_count;
    search();
    return numGoals;
}

/// bfs3d.cpp:501-547
void BFS_3D::search()
{
    int* grid = m_distance_grid.data();
    int* queue = m_queue.data();
    while (m_queue_head < m_queue_tail) {
        int currentNode = queue[m_queue_head++];
        int currentCost = grid[currentNode] + 1;
        ++expansions;
        for (int i = 0; i < 26; ++i) {
            int nn = currentNode + m_neighbor_offsets[i];
            if (grid[nn] < 0) {
                queue[m_queue_tail++] = nn;
                grid[nn] = currentCost;
            }
        }
    }
}

/// bfs3d.cpp:373-378 (undefined for out-of-bounds cells in the reference)
int BFS_3D::getDistance(int x, int y, int z) const
{
    int node = getNode(x, y, z);
    return m_distance_grid[node];
}

int BFS_3D::countWalls() const
{
    int count = 0;
    for (int i = 0; i < m_dim_xyz; ++i) {
        if (m_distance_grid[i] == WALL) {
            ++count;
        }
    }
    return count;
}

} // namespace oracle
