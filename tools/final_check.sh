set -x; SECONDS=0
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/final_tests.log 2>&1; tail -2 gpurun_out/final_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2; echo "t=$SECONDS"
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 300 gpurun_out/final_bench.err; echo "t=$SECONDS"
timeout 900 python bench.py --impl reference > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; tail -c 300 gpurun_out/final_ref.err; echo "t=$SECONDS"
python - <<EOF
import json
d=json.load(open("gpurun_out/final_bench.json")); r=json.load(open("gpurun_out/final_ref.json"))
print({k:d[k] for k in ("metric","value","unit","ms_per_step","gpu_launches","vs_baseline","dtype")})
print("clocks", d["clocks"])
e=d["e2e"]; print("e2e", e["value"], "single", e["single_context_value"], "plan", e["plan_queries_per_s"], "ubr1", e["ubr1_queries_per_s"], "dropin", e["dropin_expansions_per_s"], "dual", e["dual_arm_states_per_s"], "bfs", e["bfs_ms"], e["bfs_ms_150"])
print("roofline", {k:d["roofline"][k] for k in ("frac","achieved","peak","traffic","bfs_frac","l2_frac","states_frac")})
print("cpu", {k:d["cpu_baseline"][k] for k in ("value","cores","kind")})
print("ref arm", {k:r[k] for k in ("impl","metric","value","unit")}, r["cpu_baseline"]["cores"], r["config"]==d["config"])
EOF
