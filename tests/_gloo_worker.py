"""Worker of tests/test_host_logic.py::test_sharded_evaluator_gloo (world_size 2, CPU, gloo)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch.distributed as dist
    from nauyaca_b200.dist import ShardedEvaluator, shard_bounds
    from oracle import oracle_py
    rank, world, port, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    spec = json.load(open(os.path.join(ROOT, "tests", "golden", "systems.json")))["doc2"]
    osys = oracle_py.OracleSystem(spec)
    X = np.random.default_rng(77).random((37, spec["ndim"]))      # ragged: 19 + 18
    X[5, 0] = 2.0                                                   # one out-of-bounds row
    calls = []

    def evaluate(Xl):
        calls.append(Xl.shape[0])
        return osys.loglike(Xl, mode=0)

    ev = ShardedEvaluator(evaluate)
    chi2, logl, st = ev(X)
    lo, hi = shard_bounds(37, world, rank)
    assert calls == [hi - lo]
    np.savez(out, chi2=chi2, logl=logl, st=st)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
