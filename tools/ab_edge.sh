for b in 1 0 1 0; do SMPLGPU_V32_VERBOSE=1 SMPLGPU_V32_EDGE_BATCH=$b timeout 300 python bench.py --steps 5 --warmup 3 --plan-queries 0 --ubr1-queries 0 --dual-states 0 --no-cpu --bfs-n 0 --post-paths 0 --no-ingest --no-dropin > gpurun_out/r2ff_$b.json 2>gpurun_out/r2ff_$b.err; grep -m1 "edge kernels" gpurun_out/r2ff_$b.err; python - <<EOF
import json
d=json.load(open("gpurun_out/r2ff_$b.json"))
print("batch $b", round(d["value"]/1e9,3), d["ms_per_step"], d["roofline"]["edges_ms"], d["roofline"]["states_ms"], round(d["e2e"]["value"]/1e9,2))
EOF
done
timeout 600 python -m pytest tests/test_gpu_validity.py tests/test_gpu_dropin.py -x -q -s 2>&1 | grep -E "passed|failed|LazyARAStar|unchanged reference"
