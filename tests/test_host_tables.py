"""The host-side model builder (smpl_b200/host/robot_tables.cpp: sphere trees, motion-radius weights, checked tree
pairs -- what smplgpu_set_robot uploads) against the oracle's restatement, the reference build and the reference
build's committed golden tables, DIRECTLY (the GPU parity tests only see these tables through the verdicts they lead
to).  Reference: base_collision_models.cpp:184-222, 337-444, 569-641 (trees); robot_motion_collision_model.cpp:41-275
(weights); self_collision_model.cpp:1233-1345 (pairs)."""
import os

import numpy as np
import pytest

from conftest import ROOT
from helpers import make_oracle
from oracle_api import ref_collision_lib
from smpl_b200 import api, scenes
from test_oracle_collision import case_scene, make_reference

GOLD = os.path.join(ROOT, "tests", "golden", "collision_reference.npz")


def same_nodes(a, b):
    """Node tables agree: centre, radius, children, tree-relative ids (columns 0-6) bit for bit; the link column is an
    index into each side's own link table (the product keeps only the links a planning variable moves), so it must be
    a one-to-one relabelling."""
    if a.shape != b.shape or not np.array_equal(a[:, :7], b[:, :7]):
        return False
    fwd, back = {}, {}
    for x, y in zip(a[:, 7], b[:, 7]):
        if fwd.setdefault(x, y) != y or back.setdefault(y, x) != x:
            return False
    return True


def product_tables(scene):
    t = api.build_tables(scene)
    w, ty = t.motion_weights()
    return t, t.node_table(), w, ty, t.pairs()


@pytest.mark.parametrize("name", ["pr2_tabletop", "pr2_clutter", "ubr1_plain", "ubr1_attached_spheres", "pr2_dual_arm_15dof"])
def test_tables_equal_the_oracle(name):
    if name == "ubr1_plain":
        scene = scenes.ubr1_tabletop_scene(attach=False)
    elif name == "ubr1_attached_spheres":
        scene = scenes.ubr1_tabletop_scene(attach=True)      # attached-body tree rides behind the robot's trees
    else:
        scene, _ = case_scene(name)
    o = make_oracle(scene, with_kdl=False)
    t, nodes, w, ty, pairs = product_tables(scene)
    assert same_nodes(nodes, o.node_table())
    ow, oty = o.motion_weights()
    # joint kinds: SMPLGPU_VAR_REVOLUTE / CONTINUOUS / PRISMATIC = 0 / 1 / 2; the oracle keeps urdf's 1 / 3 / 2
    assert np.array_equal(w, ow) and np.array_equal(ty, np.array([{1: 0, 3: 1, 2: 2}[int(k)] for k in oty]))
    assert np.array_equal(pairs, o.checked_pairs())
    assert len(nodes) > 10 and (len(pairs) > 0 or scene.dof < 7)


@pytest.mark.skipif(ref_collision_lib() is None, reason="oracle/_ref/libref_collision.so not built")
@pytest.mark.parametrize("name", ["pr2_tabletop", "ubr1_plain", "pr2_dual_arm_15dof"])
def test_tables_equal_the_reference_build(name):
    scene = scenes.ubr1_tabletop_scene(attach=False) if name == "ubr1_plain" else case_scene(name)[0]
    r = make_reference(scene, None)
    t, nodes, w, ty, pairs = product_tables(scene)
    assert same_nodes(nodes, r.node_table())
    assert np.array_equal(w, r.motion_weights())
    lo, hi, cont = t.limits()
    rlo, rhi, rcont = r.limits()
    assert np.array_equal(lo, rlo) and np.array_equal(hi, rhi) and np.array_equal(np.asarray(cont, np.uint8), rcont)


@pytest.mark.parametrize("name", ["pr2_tabletop", "pr2_clutter", "pr2_dual_arm_15dof"])
def test_tables_equal_the_reference_golden(name):
    g = np.load(GOLD)
    scene, attach = case_scene(name)
    assert attach is None
    _, nodes, _, _, _ = product_tables(scene)
    assert same_nodes(nodes, g[name + "/node_table"])
