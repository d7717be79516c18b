import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from smpl_b200 import api, scenes
scene = scenes.pr2_tabletop_scene()
ctx, tables = api.setup_context(scene)
pp = scenes.PlanParams(7); pp.max_expansions = 2000
st, g = scenes.tabletop_queries(512, seed=13)
a, sa = api.plan_batch(ctx, scene, tables, pp, st, g, max_concurrent=512)
ctxs = [ctx] + [api.clone_context(ctx, scene, tables) for _ in range(5)]
b, sb = api.plan_batch(ctxs, scene, tables, pp, st, g, max_concurrent=86)
c, sc = api.plan_batch(ctxs, scene, tables, pp, st, g, max_concurrent=20)
def summ(r): return [(x["success"], x["expansions"], x["cost"], x["num_states"]) for x in r]
print("single vs multi86:", summ(a) == summ(b), " single vs multi20:", summ(a) == summ(c))
print(sa["edges_submitted"], sb["edges_submitted"], sc["edges_submitted"], sum(x["expansions"] for x in a), sum(x["expansions"] for x in b))
bad = [i for i,(x,y) in enumerate(zip(summ(a), summ(b))) if x != y]
print("mismatches", bad[:10], [ (summ(a)[i], summ(b)[i]) for i in bad[:3]])
