"""Worker of tests/test_gpu_api.py::test_device_sampler_sharded*: two ranks run the SAME
device-resident PT sampler with the proposals of every half-step split between them and the
proposal log-likelihoods all-gathered. argv: rank world port out p0.npy [backend [seed]];
backend "gloo" (both ranks on the box's GPU 0, gather staged through the host) or "nccl" (rank r
on GPU r, in-place all-gather on the device buffer); seed "none" = drawn on rank 0 and broadcast."""
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
    backend = sys.argv[6] if len(sys.argv) > 6 else "gloo"
    seed = None if (len(sys.argv) > 7 and sys.argv[7] == "none") else 1234
    device = rank if backend == "nccl" else 0
    if backend == "nccl":
        import torch
        torch.cuda.set_device(device)
    dist.init_process_group(backend, init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    spec = json.load(open(os.path.join(ROOT, "tests", "golden", "systems.json")))["doc2"]
    p0 = np.load(sys.argv[5])
    PS = nb.SpecSystem(spec)
    dev = nb.mcmc.make_device_sampler(PS, p0, tmax=50.0, seed=seed, shard=True, device=device)
    next(dev.sample(p0, iterations=1, adapt=True, storechain=False))
    p, logpost, logl = dev.advance(24, adapt=True)
    np.savez(out, p=p, logl=logl, betas=dev.betas, nacc=dev.nprop_accepted, launches=dev.engine.launch_count())
    dev.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
