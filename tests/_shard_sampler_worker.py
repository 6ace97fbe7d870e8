"""Worker of tests/test_gpu_api.py::test_device_sampler_sharded: two ranks (gloo; both on the
test box's one GPU) run the SAME device-resident PT sampler with the proposals of every
half-step split between them and the proposal log-likelihoods all-gathered."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch.distributed as dist
    import nauyaca_b200 as nb
    rank, world, port, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    spec = json.load(open(os.path.join(ROOT, "tests", "golden", "systems.json")))["doc2"]
    p0 = np.load(sys.argv[5])
    PS = nb.SpecSystem(spec)
    dev = nb.mcmc.make_device_sampler(PS, p0, tmax=50.0, seed=1234, shard=True)
    next(dev.sample(p0, iterations=1, adapt=True, storechain=False))
    p, logpost, logl = dev.advance(24, adapt=True)
    np.savez(out, p=p, logl=logl, betas=dev.betas, nacc=dev.nprop_accepted, launches=dev.engine.launch_count())
    dev.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
