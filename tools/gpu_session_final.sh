#!/bin/bash
# Round-end evidence on ONE GPU: the whole GPU suite, the default bench line, and `ncu --set full` captures of the two
# dominant kernels (cfg2 round 1 = sc_eval_fact_kernel<2,true>; msm_accumulate_kernel of a cfg3 proof). gpurun_out/$1/
R=${1:-r02z}; O=gpurun_out/$R; mkdir -p $O
export B200_PEER_TIMEOUT_S=${B200_PEER_TIMEOUT_S:-3}
timeout 400 python -m pytest tests -m gpu -q --timeout 200 -p no:cacheprovider > $O/pytest_gpu.log 2>&1
echo "suite rc=$?"; tail -3 $O/pytest_gpu.log | cut -c1-300
timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench_1gpu.json 2> $O/bench_1gpu.err
echo "bench rc=$?"; python - <<PY
import json
for line in open("$O/bench_1gpu.json"):
    if line.startswith("{"):
        d = json.loads(line)
        print("BENCH", d["value"], d["e2e"]["value"], d["phases_ms"], d["parity"]["bytes_equal"], d["sumcheck"]["ms_per_proof"],
              d["roofline"]["launch_ms"], d["roofline"]["frac"], {k: v["ms_device"] for k, v in d["legs"].items()}, d["zero_check"]["ms_per_proof"])
PY
FULL="ncu --set full --clock-control none --import-source on -f"
timeout 200 $FULL -k regex:sc_eval_fact_kernel -s 1 -c 1 -o $O/full_sc_fact_round1 python tools/prof_sumcheck.py 20 1 > /dev/null 2>&1
timeout 200 $FULL -k regex:msm_accumulate -s 0 -c 1 -o $O/full_msm_acc python tools/prof_lasso.py 20 1 > /dev/null 2>&1
for f in full_sc_fact_round1 full_msm_acc; do python tools/ncu_summary.py $O/$f.ncu-rep > $O/ncu_$f.txt 2>&1; head -12 $O/ncu_$f.txt | cut -c1-160; done
ls -la $O | head -20
