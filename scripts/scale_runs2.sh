#!/bin/bash
# Multi-GPU measurements of the final round-2 build on one 8-GPU box (see scale_runs.sh for the first set):
# C3 one ensemble sharded over 8 / 2 GPUs vs replicas, the fixed 10^7-point C5 sweep over 1/2/4/8, C4 weak at 8.
tag=${1:-r2x}
out=gpurun_out
run() { # N, name, args...
  local n=$1 name=$2; shift 2
  if [ "$n" = 1 ]; then python bench.py --gpus 1 "$@" > $out/${tag}_${name}_n1.json 2> $out/${tag}_${name}_n1.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
         bench.py --gpus $n "$@" > $out/${tag}_${name}_n$n.json 2> $out/${tag}_${name}_n$n.err; fi
  echo "$name n=$n rc=$? $(grep -h '^{' $out/${tag}_${name}_n$n.json | tail -1 | cut -c1-170)"
}
nvidia-smi -L | head -8
run 8 c3shard --workload c3 --shard-walkers --steps 200 --warmup 5
run 2 c3shard --workload c3 --shard-walkers --steps 200 --warmup 5
run 8 c3rep --workload c3 --steps 200 --warmup 5
run 1 c3 --workload c3 --steps 200 --warmup 5
for n in 8 4 2 1; do run $n c5strong --workload c5 --strong --steps 2 --warmup 1; done
run 8 c4 --workload c4 --steps 5 --warmup 3 --no-cpu
