#!/bin/bash
# Multi-GPU session: N = $2 ranks. Parity of the sharded provers over CUDA IPC / NVLink, then the strong-scaling bench.
R=${1:-r02m}; N=${2:-2}; O=gpurun_out/$R; mkdir -p $O
export B200_PEER_TIMEOUT_S=${B200_PEER_TIMEOUT_S:-5}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29511 tests/dist_worker.py > $O/dist_worker_${N}gpu.log 2>&1
echo "dist_worker rc=$?"; grep -E "OK|Error|rror:|assert|differs" $O/dist_worker_${N}gpu.log | head -10
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps ${3:-5} --warmup 3 > $O/bench_${N}gpu.json 2> $O/bench_${N}gpu.err
echo "bench rc=$?"; tail -c 2500 $O/bench_${N}gpu.json; tail -5 $O/bench_${N}gpu.err
for PROTO in ${PROTOS:-1}; do
  B200_PEER_PROTO=$PROTO timeout 300 $TR --master-port 2951$((3 + PROTO)) tools/micro/shard_rounds.py 14 18 > $O/shard_rounds_${N}gpu_proto$PROTO.log 2>&1
  echo "proto $PROTO:"; grep SHARD_ROUNDS $O/shard_rounds_${N}gpu_proto$PROTO.log
done
