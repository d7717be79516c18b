#!/bin/bash
# e2e throughput of bench.py for several host pipeline chunk sizes (SMPLGPU_HOST_CHUNK)
for c in 196608 262144 349526 524288 1048576; do
  SMPLGPU_HOST_CHUNK=$c python bench.py --steps 5 --warmup 3 --no-cpu --bfs-n 0 --plan-queries 0 2>/dev/null > /tmp/b_$c.json
  python -c "import json; d=json.load(open('/tmp/b_$c.json')); print($c, d['e2e']['value'], d['value'])"
done
