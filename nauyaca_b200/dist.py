"""Multi-GPU sharding of a walker batch.

Every parameter vector is an independent integration, so the batch shards by contiguous
row ranges with no data-path collective; the ONE exchange step is the all-gather of the
log-likelihoods so that every rank can run the sampler's accept / temperature-swap step.
One process per GPU (torch.distributed; backend "nccl" on GPUs, "gloo" in CPU tests).
"""
import numpy as np


def shard_bounds(B, world, rank):
    """Contiguous slice [lo, hi) of B rows owned by `rank` (sizes differ by at most 1)."""
    base, rem = divmod(int(B), int(world))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


class ShardedEvaluator:
    """Evaluate a replicated global batch X[B, ndim]: each rank computes its shard, then all
    ranks all-gather chi2 / logl / status so each holds the full result.

    evaluate : callable(X_local[b, ndim]) -> (chi2[b], logl[b], status[b]) as numpy arrays or
               as CUDA torch tensors (then the gather runs over NCCL without a host copy).
               In production this is BatchEngine.loglike of the rank's GPU.
    """

    def __init__(self, evaluate, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.evaluate = evaluate
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def __call__(self, X):
        import torch
        B = X.shape[0]
        lo, hi = shard_bounds(B, self.world, self.rank)
        chi2, logl, st = self.evaluate(X[lo:hi])
        if self.world == 1:
            return chi2, logl, st
        is_t = torch.is_tensor(logl)
        dev = logl.device if is_t else torch.device("cpu")
        sizes = [shard_bounds(B, self.world, r) for r in range(self.world)]
        maxb = max(h - l for l, h in sizes)
        # one packed all-gather: [chi2 | logl | status] padded to the largest shard
        pack = torch.zeros(3 * maxb, dtype=torch.float64, device=dev)
        t = (lambda a: a) if is_t else (lambda a: torch.from_numpy(np.ascontiguousarray(a)))
        n = hi - lo
        pack[0:n] = t(chi2)
        pack[maxb:maxb + n] = t(logl)
        pack[2 * maxb:2 * maxb + n] = t(st).to(torch.float64)
        out = torch.empty(self.world * 3 * maxb, dtype=torch.float64, device=dev)
        self.dist.all_gather_into_tensor(out, pack, group=self.group)
        out = out.view(self.world, 3, maxb)
        c = torch.cat([out[r, 0, :h - l] for r, (l, h) in enumerate(sizes)])
        g = torch.cat([out[r, 1, :h - l] for r, (l, h) in enumerate(sizes)])
        s = torch.cat([out[r, 2, :h - l] for r, (l, h) in enumerate(sizes)]).to(torch.int32)
        if is_t:
            return c, g, s
        return c.numpy(), g.numpy(), s.numpy()
