timeout 600 python -m pytest tests/test_gpu_validity.py tests/test_gpu_reference_golden.py tests/test_gpu_planner.py -x -q 2>&1 | tail -2
for b in 1 2 3; do timeout 300 python bench.py --steps 10 --warmup 3 --plan-queries 0 --ubr1-queries 0 --no-cpu --bfs-n 0 --post-paths 0 --no-ingest --no-dropin > gpurun_out/r2ii.json 2>gpurun_out/r2ii.err; python - <<EOF
import json
d=json.load(open("gpurun_out/r2ii.json"))
print("rep $b", round(d["value"]/1e9,3), d["ms_per_step"], d["roofline"]["edges_ms"], d["roofline"]["states_ms"], round(d["e2e"]["value"]/1e9,2), round(d["e2e"]["dual_arm_states_per_s"]/1e6))
EOF
done
