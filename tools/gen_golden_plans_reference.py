#!/usr/bin/env python
"""Writes tests/golden/plans_reference.json: results of the REFERENCE's own planning stack (oracle/ref_planner_shim.cpp,
built by `make -C oracle ref` from /root/reference) for the queries of tests/test_oracle_planner_reference.py, and
tests/golden/plans_reference_lazy.json: the same queries through the reference's lazy successors (GetLazySuccs /
GetTrueCost under its in-tree LazyARAStar), each result followed by the number of GetTrueCost calls."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import test_oracle_planner_reference as T  # noqa: E402
from test_oracle_collision import make_reference  # noqa: E402


def main():
    out = {}
    for name, (scene, attach, params, starts, goals) in T.plan_cases().items():
        r = make_reference(scene, attach)
        out[name] = [T.summary(r.plan(scene, s, g, params)) for s, g in zip(starts, goals)]
        print(name, [(x[0], x[1], x[2]) for x in out[name]])
    path = os.path.join(ROOT, "tests", "golden", "plans_reference.json")
    json.dump(out, open(path, "w"))
    print("wrote", path)
    lazy = {}
    for name, (scene, attach, params, starts, goals) in T.plan_cases().items():
        r = make_reference(scene, attach)
        lazy[name] = []
        for s, g in zip(starts, goals):
            p = r.plan(scene, s, g, params, lazy=True)
            lazy[name].append(T.summary(p) + [p["evaluations"]])
        print(name, "lazy", [(x[0], x[1], x[2], x[5]) for x in lazy[name]])
    path = os.path.join(ROOT, "tests", "golden", "plans_reference_lazy.json")
    json.dump(lazy, open(path, "w"))
    print("wrote", path)


if __name__ == "__main__":
    main()
