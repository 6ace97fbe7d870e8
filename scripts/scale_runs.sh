#!/bin/bash
# Multi-GPU measurements of the two BASELINE configs that name a split (run on an 8-GPU box):
#   C5  the fixed 10^7-point sweep, strong scaling over 1/2/4/8 GPUs (host buffers, one gather)
#   C3  the 10 x 1000-walker PT ensemble: ONE ensemble sharded over N GPUs vs one replica per GPU
# plus the NCCL path of the sharded device sampler (pytest). Lines go to gpurun_out/<tag>_*.json.
tag=${1:-r2s}
out=gpurun_out
run() { # N, name, args...
  local n=$1 name=$2; shift 2
  if [ "$n" = 1 ]; then python bench.py --gpus 1 "$@" > $out/${tag}_${name}_n1.json 2> $out/${tag}_${name}_n1.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
         bench.py --gpus $n "$@" > $out/${tag}_${name}_n$n.json 2> $out/${tag}_${name}_n$n.err; fi
  echo "$name n=$n rc=$? $(cut -c1-160 $out/${tag}_${name}_n$n.json | tail -1)"
}
nvidia-smi -L | head -8
python -m pytest tests/test_gpu_api.py -m gpu -q -k "sharded" 2>&1 | tail -3
for n in 1 2 4 8; do run $n c5strong --workload c5 --strong --steps 2 --warmup 1; done
run 1 c3 --workload c3 --steps 200 --warmup 5
for n in 2 8; do run $n c3shard --workload c3 --shard-walkers --steps 200 --warmup 5; done
for n in 2 8; do run $n c3rep --workload c3 --steps 200 --warmup 5; done
run 8 c4 --workload c4 --steps 5 --warmup 3 --no-cpu
