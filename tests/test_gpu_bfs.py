"""GPU parity tests of kernel (4), the BFS_3D wavefront: bit-exact int32 distances."""
import os

import numpy as np
import pytest

from conftest import ROOT
from oracle_api import OracleBfs
from smpl_b200 import api, scenes

pytestmark = pytest.mark.gpu
WALL = 0x7FFFFFFF


@pytest.fixture(scope="module")
def ctx():
    c = api.GpuContext(0)
    yield c
    c.close()


def oracle_grid(walls, seeds, multi=False):
    nz, ny, nx = walls.shape
    b = OracleBfs(nx, ny, nz)
    b.set_walls(walls)
    if multi:
        b.run_multi(seeds)
    else:
        b.run(*seeds)
    return b.grid()


def test_golden_fixture_from_reference_build(ctx):
    g = np.load(os.path.join(ROOT, "tests", "golden", "bfs_reference_24.npz"))
    ctx.bfs_set_walls(g["walls"])
    for k, seed in enumerate(g["seeds"]):
        # a fresh wall upload per seed: the fixture used a fresh BFS_3D per seed
        ctx.bfs_set_walls(g["walls"])
        assert ctx.bfs_run([seed]) == 1
        assert np.array_equal(ctx.bfs_download(), g["dist_%d" % k])


@pytest.mark.parametrize("dims,p,seed", [((16, 16, 16), 0.3, 1), ((40, 25, 17), 0.45, 2), ((64, 64, 64), 0.2, 3),
                                         ((7, 1, 1), 0.0, 4), ((33, 32, 31), 0.6, 5), ((130, 9, 70), 0.3, 6),
                                         ((1, 1, 1), 0.0, 7), ((97, 101, 3), 0.5, 8)])
def test_random_grids(ctx, dims, p, seed):
    rng = np.random.default_rng(seed)
    nx, ny, nz = dims
    walls = (rng.random((nz, ny, nx)) < p).astype(np.uint8)
    s = (int(rng.integers(0, nx)), int(rng.integers(0, ny)), int(rng.integers(0, nz)))
    ctx.bfs_set_walls(walls)
    assert ctx.bfs_run([s]) == 1
    got = ctx.bfs_download()
    want = oracle_grid(walls, s)
    assert np.array_equal(got, want)
    # gather API
    cells = np.stack([rng.integers(-1, nx + 1, 200), rng.integers(-1, ny + 1, 200), rng.integers(-1, nz + 1, 200)], axis=1)
    d = ctx.bfs_distances(cells)
    inb = np.all((cells >= 0) & (cells < np.array([nx, ny, nz])), axis=1)
    assert (d[~inb] == -2).all()
    assert np.array_equal(d[inb], want[cells[inb, 2] + 1, cells[inb, 1] + 1, cells[inb, 0] + 1])


def test_clutter_grid_with_unreachable_regions(ctx):
    walls = scenes.bfs_clutter_walls(96, seed=11)
    walls[40:60, 40:60, 40] = 1
    walls[40:60, 40:60, 59] = 1
    walls[40:60, 40, 40:60] = 1
    walls[40:60, 59, 40:60] = 1
    walls[40, 40:60, 40:60] = 1
    walls[59, 40:60, 40:60] = 1
    walls[45:55, 45:55, 45:55] = 0     # sealed free pocket: stays UNDISCOVERED
    seed = scenes.first_free_cell(walls, (5, 5, 5))
    ctx.bfs_set_walls(walls)
    ctx.bfs_run([seed])
    got = ctx.bfs_download()
    want = oracle_grid(walls, seed)
    assert np.array_equal(got, want)
    assert (got[46:56, 46:56, 46:56] == -1).all()
    assert ctx.bfs_last_levels() >= int(want[(want != WALL)].max())


def test_seed_quirks(ctx):
    rng = np.random.default_rng(9)
    walls = (rng.random((20, 20, 20)) < 0.3).astype(np.uint8)
    walls[10, 10, 10] = 1
    nz, ny, nx = walls.shape
    b = OracleBfs(nx, ny, nz)
    b.set_walls(walls)
    ctx.bfs_set_walls(walls)
    # out-of-bounds seed: grid reset, nothing searched, returns 0 (bfs3d.cpp:169-171)
    assert ctx.bfs_run([(25, 0, 0)]) == 0 == b.run(25, 0, 0)
    assert np.array_equal(ctx.bfs_download(), b.grid())
    # seeding a wall cell turns it into a free cell, and it stays free in later runs (bfs3d.cpp:181-187)
    assert ctx.bfs_run([(10, 10, 10)]) == 1 == b.run(10, 10, 10)
    assert np.array_equal(ctx.bfs_download(), b.grid())
    free = scenes.first_free_cell(walls, (0, 0, 0))
    ctx.bfs_run([free])
    b.run(*free)
    got = ctx.bfs_download()
    assert np.array_equal(got, b.grid())
    assert got[11, 11, 11] != WALL
    # several seeds at once (the kernel takes explicit seeds; the reference's iterator quirk that drops
    # the last triple lives in the adapter, bfs3d.h:157-211)
    seeds = [(1, 1, 1), (18, 18, 18), (5, 17, 3)]
    ctx.bfs_set_walls(walls)
    assert ctx.bfs_run(seeds) == 3
    b2 = OracleBfs(nx, ny, nz)
    b2.set_walls(walls)
    b2.run_multi(seeds + [(0, 0, 0)])   # sentinel triple so that all three real seeds are committed
    assert np.array_equal(ctx.bfs_download(), b2.grid())


def test_rerun_is_idempotent(ctx):
    walls = scenes.bfs_clutter_walls(64, seed=3)
    seed = scenes.first_free_cell(walls, (32, 32, 32))
    ctx.bfs_set_walls(walls)
    ctx.bfs_run([seed])
    a = ctx.bfs_download()
    ctx.bfs_run([seed])
    assert np.array_equal(a, ctx.bfs_download())


def test_full_size_400_cubed_properties(ctx):
    """BASELINE config[2] size.  The oracle needs ~2 s here, so compare in full as well."""
    n = 400
    walls = scenes.bfs_clutter_walls(n, seed=11)
    seed = scenes.first_free_cell(walls, (200, 200, 200))
    ctx.bfs_set_walls(walls)
    ctx.bfs_run([seed])
    got = ctx.bfs_download()
    inner = got[1:-1, 1:-1, 1:-1]
    # size-independent properties: walls stay walls, seed is 0, every discovered cell d>0 has a
    # 26-neighbour at d-1 and none below d-1 (1-Lipschitz in the Chebyshev metric)
    assert ((inner == WALL) == (walls != 0)).all()
    assert got[seed[2] + 1, seed[1] + 1, seed[0] + 1] == 0
    disc = (got >= 0) & (got != WALL)
    big = np.where(disc, got, np.int32(1 << 30))
    nmin = np.full_like(big, 1 << 30)
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                if dz or dy or dx:
                    nmin = np.minimum(nmin, np.roll(big, (dz, dy, dx), axis=(0, 1, 2)))
    pos = disc & (got > 0)
    assert (nmin[pos] == got[pos] - 1).all()
    want = oracle_grid(walls, seed)
    assert np.array_equal(got, want)
    print("400^3: %d levels, %.1f%% walls, %.1f%% discovered" % (
        ctx.bfs_last_levels(), 100.0 * (walls != 0).mean(), 100.0 * disc.mean()))


@pytest.mark.parametrize("mode", [0, 1])
def test_both_kernels_on_odd_shapes_and_quirks(ctx, mode):
    """Non-multiple-of-16 dims, a seed on a wall, a seed next to the border, multi-seed: both kernels."""
    ctx.bfs_set_mode(mode)
    rng = np.random.default_rng(77)
    for dims in ((37, 21, 50), (5, 70, 9), (33, 33, 33)):
        nx, ny, nz = dims
        walls = (rng.random((nz, ny, nx)) < 0.35).astype(np.uint8)
        seeds = [(0, 0, 0), (nx - 1, ny - 1, nz - 1), (nx // 2, ny // 2, nz // 2)]
        walls[nz // 2, ny // 2, nx // 2] = 1                      # seeding a wall cell un-walls it
        ctx.bfs_set_walls(walls)
        ctx.bfs_run(seeds)
        # (the reference's multi-seed iterator drops the last triple, bfs3d.h:157-211: feed it a sentinel)
        assert np.array_equal(ctx.bfs_download(), oracle_grid(walls, seeds + [(0, 0, 0)], multi=True)), (dims, mode)
    ctx.bfs_set_mode(ctx.BFS_AUTO)


def test_level_kernel_and_tile_kernel_agree(ctx):
    """Both wavefront kernels (one level per grid barrier / eight levels per barrier on shared-memory tiles)
    produce the reference's distances."""
    walls = scenes.bfs_clutter_walls(96, seed=21)
    seed = scenes.first_free_cell(walls, (48, 48, 48))
    ob = OracleBfs(96, 96, 96)
    ob.set_walls(walls)
    ob.run(*seed)
    want = ob.grid()
    for mode in (ctx.BFS_LEVELS, ctx.BFS_TILES):
        ctx.bfs_set_mode(mode)
        ctx.bfs_set_walls(walls)
        ctx.bfs_run([seed])
        assert np.array_equal(ctx.bfs_download(), want), mode
    ctx.bfs_set_mode(ctx.BFS_AUTO)
