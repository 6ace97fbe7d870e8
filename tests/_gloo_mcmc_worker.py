"""Worker of tests/test_host_logic.py::test_sharded_mcmc_gloo: every rank runs the SAME PT sampler
(same seed, replicated walkers); each likelihood block is sharded over the ranks and
all-gathered. CPU / gloo, the oracle stands in for the GPU engine."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch.distributed as dist
    import nauyaca_b200 as nb
    from nauyaca_b200.dist import ShardedEvaluator
    from oracle import oracle_py
    rank, world, port, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    spec = json.load(open(os.path.join(ROOT, "tests", "golden", "systems.json")))["doc2"]
    PS = nb.SpecSystem(spec)
    osys = oracle_py.OracleSystem(spec)
    nloc = []

    def local(X):
        nloc.append(X.shape[0])
        return osys.loglike(X, mode=0)
    sharded = ShardedEvaluator(local)
    rng = np.random.mtrand.RandomState(21)
    p0 = np.clip(0.5 + 0.05 * rng.normal(size=(3, 24, spec["ndim"])), 0, 1)
    s = nb.mcmc.make_sampler(PS, p0, tmax=20.0, random=np.random.mtrand.RandomState(4),
                             evaluate=lambda X: sharded(X)[1])
    p, logpost, logl = s.run_mcmc(p0, iterations=6, adapt=True)
    np.savez(out, p=p, logl=logl, nloc=np.array(nloc))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
