#!/bin/bash
# tests + default bench line on one GPU (no profiler). gpurun_out/$1/
R=${1:-r02y}; O=gpurun_out/$R; mkdir -p $O
export B200_PEER_TIMEOUT_S=${B200_PEER_TIMEOUT_S:-3}
timeout 400 python -m pytest tests -m gpu -q --timeout 200 -p no:cacheprovider ${PYTEST_ARGS:-} > $O/pytest_gpu.log 2>&1
echo "suite rc=$?"; tail -25 $O/pytest_gpu.log | cut -c1-300
timeout 600 python bench.py --steps 5 --warmup 3 ${BENCH_ARGS:-} > $O/bench_1gpu.json 2> $O/bench_1gpu.err
echo "bench rc=$?"; tail -3 $O/bench_1gpu.err; python - <<PY
import json
for line in open("$O/bench_1gpu.json"):
    if line.startswith("{"):
        d = json.loads(line)
        print("BENCH", d["value"], d["e2e"]["value"], d["phases_ms"], d["parity"]["bytes_equal"], d["gpu_launches"],
              d["sumcheck"] and d["sumcheck"]["ms_per_proof"], d["roofline"].get("launch_ms"),
              d["legs"] and {k: v["ms_device"] for k, v in d["legs"].items()}, d["roofline_msm"] and d["roofline_msm"]["frac"])
PY
