#!/bin/bash
# Re-measurement of the C5 strong-scaling sweep only (see scripts/scale_runs.sh).
tag=${1:-r2t}; out=gpurun_out
for n in 1 2 4 8; do
  if [ "$n" = 1 ]; then python bench.py --gpus 1 --workload c5 --strong --steps 2 --warmup 1 2> $out/${tag}_c5strong_n1.err | grep "^{" > $out/${tag}_c5strong_n1.json
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --workload c5 --strong --steps 2 --warmup 1 2> $out/${tag}_c5strong_n$n.err | grep "^{" > $out/${tag}_c5strong_n$n.json; fi
  echo "c5strong n=$n $(cut -c1-150 $out/${tag}_c5strong_n$n.json)"
done
