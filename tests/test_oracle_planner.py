"""The oracle's query-level driver (ManipLattice + ARA*, oracle/lattice.cpp) against frozen plans."""
import json
import os

import numpy as np

from conftest import ROOT
from helpers import make_oracle
from smpl_b200 import scenes

GOLD = os.path.join(ROOT, "tests", "golden")


def test_golden_plans():
    gold = json.load(open(os.path.join(GOLD, "pr2_tabletop_plans.json")))
    scene = scenes.pr2_tabletop_scene()
    o = make_oracle(scene)
    params = scenes.PlanParams(scene.dof)
    params.max_expansions = gold["max_expansions"]
    solved = 0
    for s, g, r in zip(gold["starts"], gold["goals"], gold["results"]):
        o.heur_init(scene.inflation_radius, scene.cost_per_cell)
        p = o.plan(np.array(s), np.array(g), params)
        assert [int(p["success"]), p["expansions"], p["cost"], p["num_states"]] == r[:4]
        assert list(map(int, p["path_ids"])) == r[4]
        solved += p["success"]
        if p["success"]:
            # path sanity: starts at the start state (id 1), ends at the goal state (id 0), unit edge costs
            assert r[4][0] == 1 and r[4][-1] == 0 and p["cost"] == 1000 * (len(r[4]) - 1)
    assert solved >= len(gold["results"]) // 2


def test_plan_is_deterministic_and_replannable():
    scene = scenes.pr2_tabletop_scene()
    o = make_oracle(scene)
    params = scenes.PlanParams(scene.dof)
    params.max_expansions = 2000
    starts, goals = scenes.tabletop_queries(2, seed=3)
    res = []
    for _ in range(2):
        o.heur_init(scene.inflation_radius, scene.cost_per_cell)
        res.append(o.plan(starts[0], goals[0], params))
    assert res[0]["expansions"] == res[1]["expansions"] and np.array_equal(res[0]["path_ids"], res[1]["path_ids"])
