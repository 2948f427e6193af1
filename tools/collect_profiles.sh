#!/bin/bash
# Round evidence, collected on the GPU box in one go (results under gpurun_out/rNN/; copy what is judged to profiles/).
R=${1:-r01}
O=gpurun_out/$R
mkdir -p $O
python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
python bench.py --steps 10 --warmup 3 > $O/bench_1gpu.json 2> $O/bench_1gpu.err
python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_reference.json 2> $O/bench_reference.err
for args in "range 4 20" "range 4 22" "and 8 22" "xor 8 20"; do python tools/bench_lasso.py $args 2>/dev/null | tail -1; done > $O/bench_lasso_tables.jsonl
for n in 20 22; do python tools/bench_zero_check.py $n 2>/dev/null | tail -2; done > $O/bench_zero_check.txt
{ python tools/bench_hyperplonk.py 14 --oracle; python tools/bench_hyperplonk.py 20; python tools/bench_hyperplonk.py 14 --lookup --oracle; python tools/bench_hyperplonk.py 20 --lookup; } 2>/dev/null | grep '^{' > $O/bench_hyperplonk.jsonl
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
$NCU --log-file $O/launches_lasso_2e20.csv python tools/prof_lasso.py 20 2 > /dev/null 2>&1
$NCU --log-file $O/launches_sumcheck_n20.csv python tools/prof_sumcheck.py 20 2 > /dev/null 2>&1
$NCU --log-file $O/launches_hyperplonk_lookup_k18.csv python tools/bench_hyperplonk.py 18 --lookup > /dev/null 2>&1
FULL="ncu --set full --clock-control none --import-source on -f"
$FULL -k regex:sc_eval_round_kernel -s 1 -c 1 -o $O/full_sc_round1 python tools/prof_sumcheck.py 20 1 > /dev/null 2>&1
$FULL -k regex:msm_accumulate -s 1 -c 1 -o $O/full_msm_acc python tools/prof_lasso.py 20 1 > /dev/null 2>&1
$FULL -k regex:sc_generic_round_kernel -s 0 -c 1 -o $O/full_generic_round0 python tools/bench_zero_check.py 20 > /dev/null 2>&1
for f in full_sc_round1 full_msm_acc full_generic_round0; do python tools/ncu_summary.py $O/$f.ncu-rep > $O/ncu_$f.txt 2>&1; done
# instruction issue rates and the multiplier's own peak (build once: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/micro/pipe_rates tools/micro/pipe_rates.cu)
[ -x tools/micro/pipe_rates ] && ./tools/micro/pipe_rates > $O/pipe_rates.txt 2>&1
# >= 2 GPUs: the sharded provers (incl. b200_dist_shard_sumchecks) against the oracle, with timing of one cooperative 2^20 proof
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  DIST_QUICK=time python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_worker.py > $O/dist_quick_2gpu.log 2>&1
  grep -E "OK|COOPERATIVE|Error|assert" $O/dist_quick_2gpu.log
fi
ls -la $O
