N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2hh_n$N.json 2> gpurun_out/r2hh_n$N.err
tail -c 400 gpurun_out/r2hh_n$N.err
python - <<EOF
import json
d=json.load(open("gpurun_out/r2hh_n$N.json"))
e=d["e2e"]
print({k:d[k] for k in ("value","n_gpus","ms_per_step")})
print("e2e", e["value"], "plan", e["plan_queries_per_s"], e["plan_expansions_per_s"], "ubr1", e["ubr1_queries_per_s"], "dual", e["dual_arm_states_per_s"], "bcast", e["dual_arm_broadcast_ms"], "parity", e["plan_parity_identical"], e["plan_parity_checked"])
EOF
