#!/bin/bash
# Heartbeat experiment at N = $2 ranks: cost of a sharded sum-check round with several heartbeat settings, then the bench
R=${1:-r02hb}; N=${2:-8}; O=gpurun_out/$R; mkdir -p $O
export B200_PEER_TIMEOUT_S=5
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1"
P=29520
for HB in ${HBS:-"1,1000,1" "148,1000,0" "148,2000,1"}; do
  P=$((P + 1))
  B200_HEARTBEAT=$HB timeout 200 $TR --master-port $P tools/micro/shard_rounds.py 14 > $O/shard_rounds_${N}gpu_hb_$HB.log 2>&1
  echo "heartbeat $HB:"; grep SHARD_ROUNDS $O/shard_rounds_${N}gpu_hb_$HB.log
done
if [ -n "$BENCH_HB" ]; then
  B200_HEARTBEAT=$BENCH_HB timeout 300 $TR --master-port 29540 bench.py --gpus $N --steps 5 --warmup 3 --no-legs > $O/bench_${N}gpu_hb.json 2> $O/bench_${N}gpu_hb.err
  echo "bench rc=$?"; python - <<PY
import json
d = json.load(open("$O/bench_${N}gpu_hb.json"))
print("BENCH_HB", d["value"], d["e2e"]["value"], d["phases_ms"], d["parity"])
PY
fi
