"""The certified single-precision validity path (validity32.cuh): its verdicts must equal the all-double
kernels' and the oracle's bit for bit, its forward kinematics must stay inside the error bound that certifies
it, and the number of items the double-precision kernels had to resolve is reported."""
import numpy as np
import pytest

from helpers import make_oracle
from smpl_b200 import api, scenes

pytestmark = pytest.mark.gpu

SCENES = [scenes.pr2_clutter_scene, scenes.pr2_tabletop_scene, scenes.ubr1_tabletop_scene, scenes.pr2_dual_arm_scene]


@pytest.mark.parametrize("maker", SCENES)
def test_single_precision_fk_stays_inside_the_certified_bound(maker):
    scene = maker()
    ctx, tables = api.setup_context(scene)
    in_use, e_pos, eps_cells = ctx.certified_bounds()
    assert in_use and 0.0 < e_pos < 2e-4 and eps_cells < 0.125
    lo, hi, cont = tables.limits()
    q = scenes.random_states(4096, lo, hi, cont, seed=61)
    q[::9] *= 3.0                                   # wrapped / out-of-limit angles too (|q| up to ~9.4)
    c64 = ctx.fk_sphere_centers(q)
    c32 = ctx.fk_sphere_centers_f32(q).astype(np.float64)
    err = np.sqrt(((c64 - c32) ** 2).sum(axis=2)).max()
    print("%s: max |centre_f32 - centre_f64| = %.3g m, certified bound %.3g m (%.1fx), eps %.3g cells" % (
        maker.__name__, err, e_pos, e_pos / err, eps_cells))
    assert err * 4.0 < e_pos
    ctx.close()


@pytest.mark.parametrize("maker", SCENES)
def test_certified_equals_exact_and_counts_resolved_items(maker):
    scene = maker()
    ctx, tables = api.setup_context(scene)
    lo, hi, cont = tables.limits()
    n = 200000 if scene.dof <= 7 else 60000
    q = scenes.random_states(n, lo, hi, cont, seed=62)
    deltas = None if scene.dof == 7 else np.eye(scene.dof)[np.arange(22) % scene.dof] * np.where(np.arange(22) % 2, -0.1, 0.12)[:, None]
    q0, q1 = scenes.mprim_edges(q, deltas)
    ctx.set_precision_mode(ctx.CERTIFIED_F32)
    v32 = ctx.is_states_valid(q)
    res_states = ctx.last_f64_resolved()
    e32, c32 = ctx.is_edges_valid(q0, q1)
    res_edges = ctx.last_f64_resolved()
    ctx.set_precision_mode(ctx.EXACT_F64)
    v64 = ctx.is_states_valid(q)
    assert ctx.last_f64_resolved() == 0
    e64, c64 = ctx.is_edges_valid(q0, q1)
    print("%s: %d states (%.1f%% valid), resolved in double: %d states (%.3f%%), %d edges (%.3f%%)" % (
        maker.__name__, n, 100 * v64.mean(), res_states, 100.0 * res_states / n, res_edges, 100.0 * res_edges / n))
    assert np.array_equal(v32, v64)
    assert np.array_equal(c32, c64)
    assert np.array_equal(e32, e64)
    assert res_states < 0.05 * n and res_edges < 0.10 * n
    ctx.close()


def test_certified_path_against_oracle_with_adversarial_states():
    """States snapped so that sphere centres sit (nearly) on cell boundaries: the certified path must still
    agree with the oracle, by sending those states to the double-precision kernel."""
    scene = scenes.pr2_clutter_scene()
    o = make_oracle(scene)
    ctx, tables = api.setup_context(scene)
    lo, hi, cont = tables.limits()
    q = scenes.random_states(30000, lo, hi, cont, seed=63)
    q = np.round(q / (np.pi / 180.0)) * (np.pi / 180.0)        # lattice states: whole degrees
    q[::3, 0] = 0.0
    q[::3, 2] = 0.0                                            # planar arm poses: many centres share coordinates
    v = ctx.is_states_valid(q)
    assert np.array_equal(v, o.is_states_valid(q))
    q0, q1 = scenes.mprim_edges(q[:8000])
    e, c = ctx.is_edges_valid(q0, q1)
    eo, co = o.is_edges_valid(q0, q1)
    assert np.array_equal(c, co) and np.array_equal(e, eo)
    # non-finite and huge joint values are the double path's business
    bad = q[:64].copy()
    bad[0, 1] = np.nan
    bad[1, 3] = 1e9
    bad[2, 4] = 200.0
    assert np.array_equal(ctx.is_states_valid(bad)[1:], o.is_states_valid(bad)[1:])
    ctx.close()
