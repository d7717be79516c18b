// ORACLE (test infrastructure only -- never linked into the product path).
//
// CPU restatement of sbpl::motion::BFS_3D (smpl/include/smpl/bfs3d/bfs3d.h,
// smpl/src/bfs3d.cpp) run synchronously (the reference's background
// std::thread and its lifecycle races -- SURVEY.md section 8 defect 6 -- are
// not restated; getDistance never spins).
//
// Pinned: tests/test_oracle_bfs.py compares this restatement with the
// reference's own bfs3d.cpp compiled into oracle/_ref/libref_bfs3d.so
// (oracle/Makefile target `ref`) on random grids, cell for cell.
#ifndef ORACLE_BFS3D_H
#define ORACLE_BFS3D_H

#include <vector>

namespace oracle {

class BFS_3D
{
public:
    static const int WALL = 0x7FFFFFFF;
    static const int UNDISCOVERED = -1; // 0xFFFFFFFF

    BFS_3D(int width, int height, int length); // bfs3d.cpp:40-111
    void setWall(int x, int y, int z);         // :132-141
    bool isWall(int x, int y, int z) const;
    bool inBounds(int x, int y, int z) const;  // bfs3d.h:151-155
    int getNode(int x, int y, int z) const;    // bfs3d.h:213-220
    int run(int x, int y, int z);              // bfs3d.cpp:156-201 (search runs inline)
    int run(const int* xyz, int count);        // bfs3d.h:157-211 multi-seed
    int getDistance(int x, int y, int z) const;// bfs3d.cpp:373-378
    int countWalls() const;

    int dimX() const { return m_dim_x; }
    int dimY() const { return m_dim_y; }
    int dimZ() const { return m_dim_z; }
    const std::vector<int>& grid() const { return m_distance_grid; }
    long long expansions;

private:
    int m_dim_x, m_dim_y, m_dim_z, m_dim_xy, m_dim_xyz;
    std::vector<int> m_distance_grid;
    std::vector<int> m_queue;
    int m_queue_head, m_queue_tail;
    int m_neighbor_offsets[26];
    void search(); // bfs3d.cpp:501-547
};

} // namespace oracle

#endif
