"""Multi-GPU contact step: one process per GPU, every rank holds the whole mesh and the positions,
works on a disjoint shard of the broad-phase candidates, and the results combine with three all-reduces
(energy: sum, gradient: sum, step size: min).  The Hessian stays the rank's additive contribution.

`ShardedContactStep` is host-side plumbing over the API classes of `api.py` and `torch.distributed`; it runs
unchanged on NCCL (CUDA tensors, product library) and on gloo (CPU tensors; used by the world-size-2 tests with
the candidate shard selected through the public Candidates container instead of the device-side Morton range).
"""
import numpy as np


class ShardedContactStep:
    def __init__(self, api, mesh, rank, world, dist=None, device=None, native=True):
        """native=True: the library shards the query leaves itself (ipcb_ctx_set_shard, product only);
        native=False: candidates are built in full and rank r keeps the slice [r*n/world, (r+1)*n/world) of every
        kind through Candidates.set (works with any library that exports the host ABI)."""
        self.api, self.mesh, self.rank, self.world, self.dist, self.device, self.native = api, mesh, rank, world, dist, device, native
        if native:
            api.lib.check(api.lib.ctx_set_shard(mesh._ctx, rank, world))

    def _slice(self, pairs):
        n = len(pairs)
        return pairs[(n * self.rank) // self.world:(n * (self.rank + 1)) // self.world]

    def _shard_candidates(self, cand):
        if self.native or self.world == 1:
            return cand
        kinds = [cand.vv_candidates, cand.ev_candidates, cand.ee_candidates, cand.fv_candidates]
        out = self.api.Candidates()
        out.set(self.mesh, *[self._slice(np.asarray(k)) for k in kinds])
        return out

    def _allreduce(self, array, op):
        if self.dist is None or self.world == 1:
            return array
        import torch

        t = torch.from_numpy(np.ascontiguousarray(array, dtype=np.float64).reshape(-1).copy())
        if self.device is not None:
            t = t.to(self.device)
        self.dist.all_reduce(t, op=op)
        return t.cpu().numpy().reshape(np.shape(array))

    def step(self, V0, V1, dhat, stiffness=1.0, psd=None, dmin=0.0, min_distance=0.0, ccd=None):
        api = self.api
        cand = api.Candidates()
        cand.build(self.mesh, V0, 0.5 * (dhat + dmin))
        cand = self._shard_candidates(cand)
        coll = api.NormalCollisions()
        coll.build(cand, self.mesh, V0, dhat, dmin)
        B = api.BarrierPotential(dhat, stiffness)
        energy = B(coll, self.mesh, V0)
        grad = B.gradient(coll, self.mesh, V0)
        H = B.hessian(coll, self.mesh, V0, api.PSDProjectionMethod.CLAMP if psd is None else psd)
        swept = api.Candidates()
        swept.build(self.mesh, V0, V1, 0.5 * min_distance)
        swept = self._shard_candidates(swept)
        step = swept.compute_collision_free_stepsize(self.mesh, V0, V1, min_distance, ccd)
        if self.dist is not None and self.world > 1:
            SUM, MIN = self.dist.ReduceOp.SUM, self.dist.ReduceOp.MIN
            energy = float(self._allreduce(np.array([energy]), SUM)[0])
            grad = self._allreduce(grad, SUM)
            step = float(self._allreduce(np.array([step]), MIN)[0])
        return dict(energy=energy, gradient=grad, hessian_local=H, step=step, collisions=coll.counts())
