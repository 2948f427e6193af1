#!/bin/bash
# One GPU session: new / risky tests first (short timeouts), then the whole GPU suite, then the bench. Results under gpurun_out/$1/
R=${1:-r02}; O=gpurun_out/$R; mkdir -p $O
export B200_PEER_TIMEOUT_S=${B200_PEER_TIMEOUT_S:-3}
timeout 900 python -m pytest tests/test_gpu_sharded_local.py tests/test_gpu_sumcheck.py -m gpu -q --timeout 100 -p no:cacheprovider -x --durations=12 > $O/pytest_new.log 2>&1
echo "new tests rc=$?"; tail -5 $O/pytest_new.log | cut -c1-400
if grep -q "failed\|error" $O/pytest_new.log; then DESEL="--deselect tests/test_gpu_sharded_local.py"; else DESEL=""; fi
timeout 1500 python -m pytest tests -m gpu -q --timeout 200 -p no:cacheprovider $DESEL > $O/pytest_gpu.log 2>&1
echo "suite rc=$?"; tail -5 $O/pytest_gpu.log | cut -c1-400
if [ "$2" != "nobench" ]; then
  timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench_1gpu.json 2> $O/bench_1gpu.err
  echo "bench rc=$?"; tail -c 3000 $O/bench_1gpu.json; tail -5 $O/bench_1gpu.err
fi
if [ "$3" == "ncu" ]; then
  NCU="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
  $NCU --log-file $O/launches_and8_22.csv python tools/prof_lasso.py 22 1 and 8 > $O/prof_and.log 2>&1
  $NCU --log-file $O/launches_range4_20.csv python tools/prof_lasso.py 20 1 range 4 > $O/prof_range.log 2>&1
  $NCU --log-file $O/launches_sumcheck_n20.csv python tools/prof_sumcheck.py 20 2 > /dev/null 2>&1
fi
if [ "$4" == "local" ]; then
  timeout 300 python tools/micro/local_ranks.py 12 > $O/local_ranks.log 2>&1; grep LOCAL_RANKS $O/local_ranks.log
fi
