#!/usr/bin/env python
"""Writes tests/golden/plans_reference.json: results of the REFERENCE's own planning stack (oracle/ref_planner_shim.cpp,
built by `make -C oracle ref` from /root/reference) for the queries of tests/test_oracle_planner_reference.py."""
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


if __name__ == "__main__":
    main()
