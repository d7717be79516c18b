for rep in 1 2 3; do for cfg in "0 2" "1 2" "0 1"; do set -- $cfg; SMPLGPU_BANK_COOPERATIVE=$1 SMPLGPU_BFS_MODE=$2 timeout 300 python bench.py --steps 3 --warmup 3 --states 65536 --bfs-n 0 --post-paths 0 --no-ingest --no-dropin --dual-states 0 --no-cpu > gpurun_out/r2gg.json 2> gpurun_out/r2gg.err; python - <<EOF
import json
d=json.load(open("gpurun_out/r2gg.json"))
e=d["e2e"]
print("cooperative $1 mode $2 rep $rep", round(e["plan_queries_per_s"]), round(e["ubr1_queries_per_s"]), round(d["plan"]["rank0"]["setup_seconds"],3), round(d["ubr1_plan"]["rank0"]["setup_seconds"],3))
EOF
done; done
