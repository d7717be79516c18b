// ORACLE (test infrastructure only -- never linked into the product path).
//
// C shim around the REFERENCE's own BFS_3D class, compiled together with the
// reference's bfs3d.cpp where it lies under /root/reference (see
// oracle/Makefile target `ref`; output oracle/_ref/libref_bfs3d.so).  It is
// used to pin oracle/bfs3d.cpp and as the `kind: "reference"` CPU baseline.
//
// Built with -fno-access-control so the shim can join the search thread and
// read the distance grid directly: BFS_3D::run() writes m_running = true
// AFTER launching the thread that ends by writing m_running = false
// (bfs3d.cpp:189-199, :546), so polling isRunning() can spin forever.
#include <cstdint>
#include <chrono>
#include <thread>

#include <smpl/bfs3d/bfs3d.h>

using sbpl::motion::BFS_3D;

// The reference's std console backend (smpl/src/console/console.cpp) needs
// boost::program_options, which is not installed; bfs3d.cpp only references
// the three symbols below (its SMPL_INFO / SMPL_ERROR log lines), so the shim
// supplies silent definitions of the interface declared in
// smpl/console/detail/console_std.h:38-56.
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

static void join_search(BFS_3D* b)
{
    if (b->m_search_thread.joinable()) {
        b->m_search_thread.join();
    }
    b->m_running = false;
}

extern "C" {

BFS_3D* ref_bfs_create(int nx, int ny, int nz) { return new BFS_3D(nx, ny, nz); }

void ref_bfs_destroy(BFS_3D* b)
{
    join_search(b);
    // ~BFS_3D joins unconditionally (bfs3d.cpp:113-115); hand it a joinable thread
    b->m_search_thread = std::thread([] { });
    delete b;
}

/// walls: one byte per cell, x-fastest unpadded (index = (z*ny + y)*nx + x)
void ref_bfs_set_walls(BFS_3D* b, const uint8_t* walls)
{
    int nx, ny, nz;
    b->getDimensions(&nx, &ny, &nz);
    for (int z = 0; z < nz; ++z) {
    for (int y = 0; y < ny; ++y) {
    for (int x = 0; x < nx; ++x) {
        if (walls[((size_t)z * ny + y) * nx + x]) {
            b->setWall(x, y, z);
        }
    }
    }
    }
}

int ref_bfs_run(BFS_3D* b, int x, int y, int z)
{
    join_search(b);
    int r = b->run(x, y, z);
    join_search(b);
    return r;
}

int ref_bfs_run_multi(BFS_3D* b, const int32_t* xyz, int count)
{
    join_search(b);
    int r = b->run(xyz, xyz + 3 * count);
    join_search(b);
    return r;
}

/// seconds for run() + search to completion (thread spawn included, as a
/// caller of the reference would experience before its first blocking read)
double ref_bfs_time_run(BFS_3D* b, int x, int y, int z)
{
    join_search(b);
    auto t0 = std::chrono::steady_clock::now();
    b->run(x, y, z);
    join_search(b);
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

int ref_bfs_grid(BFS_3D* b, int32_t* out)
{
    join_search(b);
    for (int i = 0; i < b->m_dim_xyz; ++i) {
        out[i] = b->m_distance_grid[i];
    }
    return b->m_dim_xyz;
}

int ref_bfs_get_distance(BFS_3D* b, int x, int y, int z)
{
    join_search(b);
    if (!b->inBounds(x, y, z)) {
        return -2;
    }
    return b->getDistance(x, y, z);
}

int ref_bfs_count_walls(BFS_3D* b) { join_search(b); return b->countWalls(); }

} // extern "C"
