// ORACLE (test infrastructure only -- never linked into the product path).
//
// C shim around the REFERENCE's own EuclidDistanceMap (smpl/src/distance_map/euclid_distance_map.cpp +
// distance_map_common.cpp + the templates in smpl/include/smpl/distance_map/detail/distance_map.hpp,
// compiled where they lie under /root/reference by `make -C oracle ref` against the Eigen stand-in in
// oracle/ref_stubs/Eigen; output oracle/_ref/libref_distmap.so).  It pins oracle/distance_map.cpp.
//
// Fork defect 1: this fork's constructor has the interior-cell initialisation
// commented out (distance_map.hpp:166-177), so a freshly built map holds uninitialised cells.  The shim,
// compiled with -fno-access-control, performs exactly those commented-out statements (resetCell, x/y/z) on
// every interior cell and then re-runs the reference's own initBorderCells() + propagateBorder(), i.e. it
// restores the upstream constructor with the reference's code.
#include <cstdarg>
#include <sstream>
#include <string>
#include <vector>

#include <smpl/console/console.h>
#include <smpl/distance_map/euclid_distance_map.h>

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

// The fork's constructor runs initBorderCells() + propagateBorder() over interior cells it never initialised
// (Grid3::resize leaves POD memory as the allocator hands it out, detail/grid.hpp:388-396).  In a fresh
// process that memory is zero pages; inside a long-lived python process it is recycled heap and the
// propagation can walk garbage.  This library is linked with -Bsymbolic and its allocations go through
// calloc, so the reference always sees the zero-filled memory of the fresh-process case.
#include <cstdlib>
#include <new>
void* operator new(std::size_t n) { void* p = std::calloc(1, n ? n : 1); if (!p) throw std::bad_alloc(); return p; }
void* operator new[](std::size_t n) { void* p = std::calloc(1, n ? n : 1); if (!p) throw std::bad_alloc(); return p; }
void operator delete(void* p) noexcept { std::free(p); }
void operator delete[](void* p) noexcept { std::free(p); }
void operator delete(void* p, std::size_t) noexcept { std::free(p); }
void operator delete[](void* p, std::size_t) noexcept { std::free(p); }

using sbpl::EuclidDistanceMap;

extern "C" {

EuclidDistanceMap* ref_distmap_create(double ox, double oy, double oz, double sx, double sy, double sz,
                                      double res, double max_dist)
{
    EuclidDistanceMap* m = new EuclidDistanceMap(ox, oy, oz, sx, sy, sz, res, max_dist);
    // the statements the fork commented out (distance_map.hpp:171-174), then border init + propagation again
    for (int x = 1; x < m->m_cells.xsize() - 1; ++x) {
        for (int y = 1; y < m->m_cells.ysize() - 1; ++y) {
            for (int z = 1; z < m->m_cells.zsize() - 1; ++z) {
                auto& c = m->m_cells(x, y, z);
                m->resetCell(c);
                c.x = x;
                c.y = y;
                c.z = z;
            }
        }
    }
    m->initBorderCells();
    m->propagateBorder();
    return m;
}

void ref_distmap_destroy(EuclidDistanceMap* m) { delete m; }

void ref_distmap_dims(EuclidDistanceMap* m, int* dims)
{
    dims[0] = m->numCellsX();
    dims[1] = m->numCellsY();
    dims[2] = m->numCellsZ();
}

static std::vector<Eigen::Vector3d> to_points(const double* xyz, int n)
{
    std::vector<Eigen::Vector3d> pts;
    pts.reserve(n);
    for (int i = 0; i < n; ++i) {
        pts.push_back(Eigen::Vector3d(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]));
    }
    return pts;
}

void ref_distmap_add_points(EuclidDistanceMap* m, const double* xyz, int n) { m->addPointsToMap(to_points(xyz, n)); }
void ref_distmap_remove_points(EuclidDistanceMap* m, const double* xyz, int n) { m->removePointsFromMap(to_points(xyz, n)); }
void ref_distmap_update_points(EuclidDistanceMap* m, const double* old_xyz, int n_old, const double* new_xyz, int n_new)
{
    m->updatePointsInMap(to_points(old_xyz, n_old), to_points(new_xyz, n_new));
}

/// squared cell distance of every interior cell, x-major / z-fastest (the layout of oracle df_d2 and of the
/// device field)
void ref_distmap_d2(EuclidDistanceMap* m, int* out)
{
    const int nx = m->numCellsX(), ny = m->numCellsY(), nz = m->numCellsZ();
    for (int x = 0; x < nx; ++x) {
        for (int y = 0; y < ny; ++y) {
            for (int z = 0; z < nz; ++z) {
                out[((size_t)x * ny + y) * nz + z] = m->m_cells(x + 1, y + 1, z + 1).dist;
            }
        }
    }
}

/// DistanceMap::getDistance(double, double, double) (distance_map.hpp:281-300), metres
double ref_distmap_distance(EuclidDistanceMap* m, double x, double y, double z) { return m->getMetricDistance(x, y, z); }

void ref_distmap_world_to_grid(EuclidDistanceMap* m, double x, double y, double z, int* g)
{
    m->worldToGrid(x, y, z, g[0], g[1], g[2]);
}

} // extern "C"
