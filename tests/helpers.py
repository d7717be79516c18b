"""Shared test helpers: build an oracle scene and a GPU context from the same smpl_b200.scenes.Scene."""
import numpy as np

from oracle_api import OracleScene


def make_oracle(scene, prime_q=None, with_kdl=True):
    o = OracleScene(scene.robot_path, scene.group, scene.planning_joints, scene.origin, scene.size, scene.res,
                    scene.max_dist)
    for k, v in scene.fixed_joints.items():
        assert o.set_joint(k, v) == 0
    if scene.use_desc_acm:
        o.use_desc_acm()
    for a, b, allowed in scene.acm_extra:
        o.acm_set(a, b, allowed)
    if scene.padding:
        o.set_padding(scene.padding)
    if scene.attached is not None:
        body_id, link, centers, radius = scene.attached
        o.attach_spheres(body_id, link, centers, radius)
    if len(scene.cells):
        o.add_cells(scene.cells)
    if len(getattr(scene, "boxes", [])):
        o.insert_boxes(scene.boxes)     # WorldCollisionModel::insertObject: VoxelizeBox + addPointsToField
    if with_kdl and scene.chain_root is not None:
        o.init_kdl(scene.chain_root, scene.chain_tip, scene.planning_link, scene.T_kin_to_planning, scene.xyz_offset)
    o.prime(np.zeros(scene.dof) if prime_q is None else prime_q)
    return o


def flips_within_tolerance(gpu, cpu, cell_margin, pair_margin, tol=1e-5):
    """Parity accounting of BASELINE.json: verdicts bit-exact except states with a sphere within
    `tol` metres of a decision threshold; returns (n_flips, n_unexplained)."""
    flips = np.flatnonzero(gpu != cpu)
    bad = [i for i in flips if min(cell_margin[i], pair_margin[i]) > tol]
    return len(flips), len(bad)
