"""Pin the BFS oracle (oracle/bfs3d.cpp) against the REFERENCE's own BFS_3D compiled from
/root/reference (oracle/_ref/libref_bfs3d.so) and against the committed golden fixture that was
produced by that reference build (tools/gen_golden.py)."""
import os

import numpy as np
import pytest

from conftest import ROOT
from oracle_api import OracleBfs, RefBfs, ref_lib
from smpl_b200 import scenes

WALL = 0x7FFFFFFF


def random_walls(rng, nx, ny, nz, p):
    return (rng.random((nz, ny, nx)) < p).astype(np.uint8)


def brute_force_bfs(walls, seed):
    """Independent python/numpy 26-connected wavefront (dilation) for tiny grids."""
    nz, ny, nx = walls.shape
    dist = np.full((nz + 2, ny + 2, nx + 2), -1, np.int32)
    w = np.ones((nz + 2, ny + 2, nx + 2), bool)
    w[1:-1, 1:-1, 1:-1] = walls != 0
    dist[w] = WALL
    x, y, z = seed
    dist[z + 1, y + 1, x + 1] = 0
    front = np.zeros_like(w)
    front[z + 1, y + 1, x + 1] = True
    level = 0
    while front.any():
        level += 1
        d = np.zeros_like(front)
        for dz in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    d |= np.roll(front, (dz, dy, dx), axis=(0, 1, 2))
        front = d & (dist < 0)
        dist[front] = level
    return dist


def test_golden_fixture_from_reference_build():
    g = np.load(os.path.join(ROOT, "tests", "golden", "bfs_reference_24.npz"))
    walls, seeds = g["walls"], g["seeds"]
    for k, seed in enumerate(seeds):
        b = OracleBfs(walls.shape[2], walls.shape[1], walls.shape[0])
        b.set_walls(walls)
        b.run(*seed)
        assert np.array_equal(b.grid(), g["dist_%d" % k]), "oracle BFS differs from the reference build's golden output"


@pytest.mark.skipif(ref_lib() is None, reason="oracle/_ref not built (needs /root/reference at build time)")
@pytest.mark.parametrize("dims,p,seed", [((16, 16, 16), 0.3, 1), ((40, 25, 17), 0.45, 2), ((64, 64, 64), 0.2, 3),
                                         ((7, 1, 1), 0.0, 4), ((33, 32, 31), 0.6, 5)])
def test_oracle_matches_reference_build(dims, p, seed):
    rng = np.random.default_rng(seed)
    nx, ny, nz = dims
    walls = random_walls(rng, nx, ny, nz, p)
    sx, sy, sz = (int(rng.integers(0, d)) for d in dims)
    o, r = OracleBfs(nx, ny, nz), RefBfs(nx, ny, nz)
    o.set_walls(walls)
    r.set_walls(walls)
    assert o.run(sx, sy, sz) == r.run(sx, sy, sz)
    assert np.array_equal(o.grid(), r.grid())
    # second run on the same object: the seed of the first run stays un-walled (bfs3d.cpp:181-187)
    s2 = (int(rng.integers(0, nx)), int(rng.integers(0, ny)), int(rng.integers(0, nz)))
    o.run(*s2)
    r.run(*s2)
    assert np.array_equal(o.grid(), r.grid())


@pytest.mark.skipif(ref_lib() is None, reason="oracle/_ref not built")
def test_multi_seed_quirk_matches_reference():
    """bfs3d.h:157-211 commits a triple only when another element follows: the last seed is dropped."""
    rng = np.random.default_rng(7)
    walls = random_walls(rng, 20, 20, 20, 0.25)
    seeds = [(1, 2, 3), (15, 15, 15), (8, 3, 17)]
    o, r = OracleBfs(20, 20, 20), RefBfs(20, 20, 20)
    o.set_walls(walls)
    r.set_walls(walls)
    assert o.run_multi(seeds) == r.run_multi(seeds) == 2
    assert np.array_equal(o.grid(), r.grid())


def test_oracle_matches_brute_force_dilation():
    rng = np.random.default_rng(11)
    walls = random_walls(rng, 14, 12, 10, 0.35)
    seed = scenes.first_free_cell(walls, (3, 3, 3))
    b = OracleBfs(14, 12, 10)
    b.set_walls(walls)
    b.run(*seed)
    assert np.array_equal(b.grid(), brute_force_bfs(walls, seed))


def test_out_of_bounds_seed_resets_and_returns_zero():
    walls = np.zeros((5, 5, 5), np.uint8)
    b = OracleBfs(5, 5, 5)
    b.set_walls(walls)
    b.run(2, 2, 2)
    assert b.run(9, 0, 0) == 0
    g = b.grid()
    assert (g[1:-1, 1:-1, 1:-1] == -1).all()
