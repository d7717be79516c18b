// ORACLE (test infrastructure only -- never linked into the product path).
//
// CPU restatement of sbpl::EuclidDistanceMap / DistanceMap<Derived>
// (smpl/include/smpl/distance_map/detail/distance_map.hpp,
//  smpl/src/distance_map/{distance_map_common,euclid_distance_map}.cpp) and the
// thin OccupancyGrid read API on top of it (smpl/include/smpl/occupancy_grid.h).
//
// Fork defect 1: the fork's constructor no longer resets
// interior cells nor sets their x,y,z (distance_map.hpp:166-177, lines 171-174
// commented out), leaving them uninitialised.  The oracle follows the
// commented-out (= upstream) lines: resetCell(c); c.x=x; c.y=y; c.z=z.
// The fork-only `counter` instrumentation is ignored.
//
// parity unpinned: the reference has no golden vectors for this class.
#ifndef ORACLE_DISTANCE_MAP_H
#define ORACLE_DISTANCE_MAP_H

#include <array>
#include <utility>
#include <vector>

#include "omath.h"

namespace oracle {

class EuclidDistanceMap
{
public:
    EuclidDistanceMap(
        double origin_x, double origin_y, double origin_z,
        double size_x, double size_y, double size_z,
        double resolution, double max_dist);

    void addPointsToMap(const std::vector<Vec3>& points);
    void removePointsFromMap(const std::vector<Vec3>& points);
    /// add by effective grid coordinates (what addPointsToMap does after worldToGrid)
    void addCellsToMap(const std::vector<std::array<int, 3>>& cells);
    void reset();

    int numCellsX() const { return m_nx - 2; }
    int numCellsY() const { return m_ny - 2; }
    int numCellsZ() const { return m_nz - 2; }
    double resolution() const { return m_res; }
    double originX() const { return m_origin_x; }
    double originY() const { return m_origin_y; }
    double originZ() const { return m_origin_z; }
    int dmaxSqrd() const { return m_dmax_sqrd_int; }

    double getDistance(double x, double y, double z) const;  // distance_map.hpp:281-286
    double getDistance(int x, int y, int z) const;           // :292-300
    double getMetricSquaredDistance(double x, double y, double z) const // distance_map_interface.h:113-114
    { double d = getDistance(x, y, z); return d * d; }
    int getSquaredCellDistance(int x, int y, int z) const; // raw integer d^2, 0 when out of bounds
    void gridToWorld(int x, int y, int z, double& wx, double& wy, double& wz) const;
    void worldToGrid(double wx, double wy, double wz, int& x, int& y, int& z) const;
    bool isCellValid(int x, int y, int z) const;

    mutable long long lookups; // instrumentation: number of getDistance(double^3) calls

private:
    struct Cell
    {
        int x, y, z;
        int dist, dist_new;
        Cell* obs;
        int bucket;
        int dir;
        int pos;
    };

    static const int NUM_DIRECTIONS = 2 * 27;
    static const int NON_BORDER_NEIGHBOR_LIST_SIZE = 460;
    static const int BORDER_NEIGHBOR_LIST_SIZE = 316;
    static const int NEIGHBOR_LIST_SIZE = NON_BORDER_NEIGHBOR_LIST_SIZE + BORDER_NEIGHBOR_LIST_SIZE;

    double m_origin_x, m_origin_y, m_origin_z, m_size_x, m_size_y, m_size_z, m_res;
    int m_nx, m_ny, m_nz; // padded cell counts
    std::vector<Cell> m_cells; // x-major, z-fastest (detail/grid.hpp:361-366)
    double m_max_dist, m_inv_res;
    int m_dmax_int, m_dmax_sqrd_int, m_bucket;
    int m_no_update_dir;

    std::array<std::array<int, 3>, 27> m_neighbors;
    std::array<int, NEIGHBOR_LIST_SIZE> m_indices;
    std::array<std::pair<int, int>, NUM_DIRECTIONS> m_neighbor_ranges;
    std::array<int, NEIGHBOR_LIST_SIZE> m_neighbor_offsets;
    std::array<int, NEIGHBOR_LIST_SIZE> m_neighbor_dirs;
    std::vector<double> m_sqrt_table;
    std::vector<std::vector<Cell*>> m_open;
    std::vector<Cell*> m_rem_stack;

    Cell& cell(int x, int y, int z) { return m_cells[((size_t)x * m_ny + y) * m_nz + z]; }
    const Cell& cell(int x, int y, int z) const { return m_cells[((size_t)x * m_ny + y) * m_nz + z]; }

    void resetCell(Cell& c) const;
    void initBorderCells();
    void updateVertex(Cell* c);
    int distance(const Cell& n, const Cell& s) const;
    void lower(Cell* s);
    void raise(Cell* s);
    void waveout(Cell* n);
    void propagate();
    void lowerBounded(Cell* s);
    void propagateRemovals();
    void propagateBorder();
};

} // namespace oracle

#endif
