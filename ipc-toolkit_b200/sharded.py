"""Multi-GPU contact step: one process per GPU, every rank holds the whole mesh and the positions.

1. Broad phase + classification are sharded: rank r traverses only its Morton range of query leaves
   (`ipcb_ctx_set_shard`), so the ranks produce disjoint candidate shards with no exchange.
2. The per-rank collision sets are the reference's per-thread builders at rank granularity: ONE all-gather of the
   packed records (16-24 B each) and `NormalCollisionsBuilder::merge` on every rank (`collisions_append*` +
   `collisions_merge`) leave the full, canonical `NormalCollisions` on every rank — what the API returns anyway.
3. The potential is sharded over that common set: energy and gradient by collision range (sum all-reduces), the
   Hessian by row block — rank r assembles the DOF rows of its vertex range from every collision touching them and
   keeps only those rows, so the rank matrices tile the global CSR and need no collective.  Row-block boundaries are
   chosen to balance the number of 3x3 block contributions (`hessian_balanced_row_blocks`, identical on all ranks).
4. The step size is the min all-reduce of the per-shard earliest times of impact.

`ShardedContactStep` is the host-side form over the API classes of `api.py` and `torch.distributed` (works on gloo
with any library exporting the host ABI: the world-size-2 CPU tests drive it through the oracle, with the candidate
shard selected through the public Candidates container instead of the device-side Morton range).
`DeviceShardedStep` is the device-resident form used by `bench.py` and the NCCL test: device positions in, device
energy / gradient / step size out, nothing staged through the host except four counts per rank.
"""
import ctypes as C
import queue
import threading

import numpy as np


class Lane(threading.Thread):
    """A host thread that issues library calls for one context, so that two contexts (two sets of CUDA streams) can be
    driven concurrently: ctypes releases the GIL for the duration of a call, and every library call blocks its calling
    thread for its own count read-backs.  submit(fn) returns a handle whose wait() re-raises what fn raised."""

    class Job:
        def __init__(self, fn):
            self.fn, self.done, self.error = fn, threading.Event(), None

        def wait(self):
            self.done.wait()
            if self.error is not None:
                raise self.error

    def __init__(self):
        super().__init__(daemon=True)
        self.q = queue.SimpleQueue()
        self.start()

    def run(self):
        while True:
            job = self.q.get()
            if job is None:
                return
            try:
                job.fn()
            except BaseException as e:  # handed to the waiting thread
                job.error = e
            job.done.set()

    def submit(self, fn):
        job = Lane.Job(fn)
        self.q.put(job)
        return job

    def close(self):
        self.q.put(None)
        self.join()


class ShardedContactStep:
    def __init__(self, api, mesh, rank, world, dist=None, device=None, native=True, row_block=True):
        """native=True: the library shards the query leaves itself (ipcb_ctx_set_shard, product only);
        native=False: candidates are built in full and rank r keeps the slice [r*n/world, (r+1)*n/world) of every
        kind through Candidates.set (works with any library that exports the host ABI).
        row_block=False keeps the simpler combination: every rank's potential over its own collision shard, the
        Hessian as the rank's additive contribution (sum over ranks == global matrix)."""
        self.api, self.mesh, self.rank, self.world, self.dist, self.device, self.native = api, mesh, rank, world, dist, device, native
        self.row_block = row_block and world > 1 and dist is not None
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

    def _gather_collisions(self, coll, dmin, disjoint=True):
        """all ranks' records -> the merged, canonical set on every rank"""
        mine = (coll.vv_collisions, coll.ev_collisions, coll.ee_collisions, coll.fv_collisions)
        parts = [None] * self.world
        self.dist.all_gather_object(parts, [(r.ids, r.weight, r.eps_x, r.dtype) for r in mine])
        import types

        builders = [[types.SimpleNamespace(ids=k[0], weight=k[1], eps_x=k[2], dtype=k[3]) for k in p] for p in parts]
        full = self.api.NormalCollisions()
        full.assign(self.mesh, builders, dmin, disjoint_shards=disjoint)
        return full

    def step(self, V0, V1, dhat, stiffness=1.0, psd=None, dmin=0.0, min_distance=0.0, ccd=None, improved_max_approx=False,
             use_area_weighting=False):
        api = self.api
        cand = api.Candidates()
        cand.build(self.mesh, V0, 0.5 * (dhat + dmin))
        cand = self._shard_candidates(cand)
        coll = api.NormalCollisions()
        coll.set_use_area_weighting(use_area_weighting)
        improved = improved_max_approx and self.world > 1
        if improved_max_approx:
            coll.set_collision_set_type(api.NormalCollisions.CollisionSetType.IMPROVED_MAX_APPROX)
        if improved:
            # CollisionSetType::IMPROVED_MAX_APPROX over ranks: a sub-element pair can come from candidates of several ranks and
            # its corrections must be added once — exchange the pairs, every rank takes a slice of the united lists
            if not self.row_block:
                raise ValueError("IMPROVED_MAX_APPROX over ranks needs the united set (row_block=True)")
            coll.build(cand, self.mesh, V0, dhat, dmin, defer_corrections=True)
            parts = [None] * self.world
            self.dist.all_gather_object(parts, coll.correction_keys())
            coll.apply_corrections([np.concatenate([p[k] for p in parts]) for k in range(4)], self.rank, self.world)
        else:
            coll.build(cand, self.mesh, V0, dhat, dmin)
        shard_counts = coll.counts()
        rows = None
        if self.row_block:
            coll = self._gather_collisions(coll, dmin, disjoint=not improved)
            bounds = self.mesh.balanced_row_blocks(self.world)
            rows = (int(bounds[self.rank]), int(bounds[self.rank + 1]))
            self.mesh.set_collision_range(self.rank, self.world)
            self.mesh.set_row_block(*rows)
        B = api.BarrierPotential(dhat, stiffness)
        energy = B(coll, self.mesh, V0)
        grad = B.gradient(coll, self.mesh, V0)
        H = B.hessian(coll, self.mesh, V0, api.PSDProjectionMethod.CLAMP if psd is None else psd)
        if self.row_block:
            self.mesh.set_collision_range(0, 1)
            self.mesh.set_row_block()
        swept = api.Candidates()
        swept.build(self.mesh, V0, V1, 0.5 * min_distance)
        swept = self._shard_candidates(swept)
        step = swept.compute_collision_free_stepsize(self.mesh, V0, V1, min_distance, ccd)
        if self.dist is not None and self.world > 1:
            SUM, MIN = self.dist.ReduceOp.SUM, self.dist.ReduceOp.MIN
            energy = float(self._allreduce(np.array([energy]), SUM)[0])
            grad = self._allreduce(grad, SUM)
            step = float(self._allreduce(np.array([step]), MIN)[0])
        return dict(energy=energy, gradient=grad, hessian_local=H, rows=rows, step=step, shard_collisions=shard_counts,
                    collisions=coll.counts())


def packed_bytes(counts):
    """size of the exchange buffer of include/ipcb200.h (collisions_pack_dev) for counts [VV, EV, EE, FV]"""
    n0, n1, n2, n3 = (int(c) for c in counts)
    return 16 * (n0 + n1 + n3) + 24 * n2 + (n2 + 7) // 8 * 8


class DeviceShardedStep:
    """The contact step with device-resident inputs and outputs (product library; NCCL when world > 1).

    dV0 / dV1: torch CUDA tensors holding N x 3 column-major positions; d_energy (1), d_grad (3N), d_step (1): torch CUDA
    float64 outputs.  All library calls and collectives are enqueued on the context's stream.

    ccd_mesh: a SECOND context on the same mesh (same device).  The step size does not depend on the collision set, and
    its kernels (swept traversal, Tight-Inclusion search) wait on dependent gathers while the potential's kernels are
    bound by the FP64 pipe and HBM: with a second context the CCD half runs on its own streams, issued by its own host
    thread, beside the build + potential half — the reference's callers are free to do the same with two BroadPhase
    objects.  Without it the five calls run one after the other on one context."""

    def __init__(self, api, mesh, rank, world, dist, torch, stream, row_block=True, ccd_mesh=None):
        self.api, self.lib, self.mesh, self.ctx = api, api.lib, mesh, mesh._ctx
        self.rank, self.world, self.dist, self.torch, self.stream = rank, world, dist, torch, stream
        self.row_block = row_block and world > 1
        self.lib.check(self.lib.ctx_set_shard(self.ctx, rank, world))
        self.counts = (C.c_int64 * 4)()
        self.nnz = C.c_int64()
        self.send = self.recv = None
        self.cap = 0
        self.h_counts = torch.zeros(4, dtype=torch.int64).pin_memory()
        self.h_headers = None
        self.rows = (0, mesh.num_vertices())
        self.shard_counts = [0, 0, 0, 0]
        self.after = None  # optional callback(ctx) after every library call (bench.py collects stage times)
        self.improved = False
        self.side = self.ev_packed = self.ev_gathered = None
        if self.row_block:
            self.side = torch.cuda.Stream()
            self.ev_packed, self.ev_gathered = torch.cuda.Event(), torch.cuda.Event()
        self.ccd_mesh, self.ctx_b, self.lane, self.stream_b = ccd_mesh, None, None, None
        if ccd_mesh is not None:
            self.ctx_b = ccd_mesh._ctx
            self.lib.check(self.lib.ctx_set_shard(self.ctx_b, rank, world))
            self.stream_b = torch.cuda.ExternalStream(self.lib.ctx_stream(self.ctx_b), device=torch.device("cuda", torch.cuda.current_device()))
            self.ev_start, self.ev_ccd = torch.cuda.Event(), torch.cuda.Event()
            self.lane = Lane()

    def _done(self, ctx=None):
        if self.after is not None:
            self.after(self.ctx if ctx is None else ctx)

    def release(self):
        """drop every torch object that lives on the contexts' streams and stop the second lane (call before the meshes /
        contexts are destroyed)"""
        if self.lane is not None:
            self.lane.close()
            self.lane = None
        self.send = self.recv = self.h_counts = self.h_headers = None
        self.side = self.ev_packed = self.ev_gathered = None
        self.stream = self.stream_b = None

    HEADER = 32  # bytes in front of every rank's slot: its four record counts (int64)

    def _resize_exchange(self, need):
        torch = self.torch
        self.cap = int(need * 1.25) + 4096
        self.cap -= self.cap % 16
        self.send = torch.empty(self.cap, dtype=torch.uint8, device="cuda")
        self.recv = torch.empty(self.cap * self.world, dtype=torch.uint8, device="cuda")
        self.h_headers = torch.zeros(self.world * 4, dtype=torch.int64).pin_memory()

    def _gather_once(self):
        """[header | packed records] of every rank in ONE all-gather of `cap` bytes per rank on the side stream.  The slot
        size was agreed on in an earlier step (every rank derives it from the same gathered counts), so nothing has to be
        exchanged — or waited for — before the records travel; a rank whose records do not fit sends its header only and
        the exchange is repeated with a larger slot (every rank sees that in the headers)."""
        torch, lib, dist = self.torch, self.lib, self.dist
        need = self.HEADER + packed_bytes(self.shard_counts)
        self.h_counts.copy_(torch.tensor(self.shard_counts, dtype=torch.int64))
        with torch.cuda.stream(self.stream):
            self.send[:self.HEADER].view(torch.int64).copy_(self.h_counts, non_blocking=True)
        if need <= self.cap and sum(self.shard_counts) > 0:
            nbytes = C.c_int64()
            lib.check(lib.collisions_pack_dev(self.ctx, C.c_void_p(self.send.data_ptr() + self.HEADER), self.cap - self.HEADER, C.byref(nbytes)))
        self.ev_packed.record(self.stream)
        with torch.cuda.stream(self.side):
            self.side.wait_event(self.ev_packed)
            dist.all_gather_into_tensor(self.recv, self.send)
            self.h_headers.view(self.world, 4).copy_(self.recv.view(self.world, self.cap)[:, :self.HEADER].contiguous().view(torch.int64).view(self.world, 4),
                                                     non_blocking=True)
            self.ev_gathered.record(self.side)

    def start_exchange(self):
        """pack and all-gather the rank's records on a side stream: it overlaps with whatever runs next (the CCD half)"""
        self.shard_counts = list(self.counts)
        if self.cap == 0:  # first step: a generous slot from this rank's own records (ranks hold similar shares)
            self._resize_exchange(2 * (self.HEADER + packed_bytes(self.shard_counts)) + (1 << 20))
            # the ranks must agree on the slot size: one MAX all-reduce, first step only
            t = self.torch.tensor([self.cap], dtype=self.torch.int64, device="cuda")
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            self._resize_exchange(int(t.item()))
        self._gather_once()

    def finish_exchange(self, dmin):
        """NormalCollisionsBuilder::merge over the ranks' records, then the balanced row blocks"""
        lib = self.lib
        for attempt in range(3):
            self.ev_gathered.synchronize()  # the headers are on the host now
            self.all_counts = self.h_headers.view(self.world, 4).tolist()
            need = max(self.HEADER + packed_bytes(c) for c in self.all_counts)
            if need <= self.cap:
                break
            self._resize_exchange(need)  # identical decision on every rank: repeat the exchange with room
            self._gather_once()
        else:
            raise RuntimeError("collision exchange: slot overflow persisted")
        self.stream.wait_event(self.ev_gathered)
        lib.check(lib.collisions_clear(self.ctx))
        for r in range(self.world):
            if sum(self.all_counts[r]) == 0:
                continue
            c = (C.c_int64 * 4)(*self.all_counts[r])
            lib.check(lib.collisions_append_packed_dev(self.ctx, C.c_void_p(self.recv.data_ptr() + r * self.cap + self.HEADER), c))
        # 1 = IPCB_MERGE_DISJOINT_SHARDS; IMPROVED_MAX_APPROX correction records of different ranks can coincide: full merge
        lib.check(lib.collisions_merge(self.ctx, dmin, 0 if self.improved else 1, self.counts))
        self._done()
        bounds = self.mesh.balanced_row_blocks(self.world)
        self.rows = (int(bounds[self.rank]), int(bounds[self.rank + 1]))

    def exchange_correction_keys(self):
        """CollisionSetType::IMPROVED_MAX_APPROX over ranks (include/ipcb200.h, collisions_corrections_*): the build stopped
        after this rank's unique sub-element keys; all-gather the four key lists, let every rank unite them and add the
        corrections of its slice.  Sizes first (one small all-gather), then one padded all-gather of the keys."""
        torch, lib, dist = self.torch, self.lib, self.dist
        n = (C.c_int64 * 4)()
        lib.check(lib.collisions_corrections_keys_dev(self.ctx, n))
        mine = torch.tensor(list(n), dtype=torch.int64, device="cuda")
        every = torch.empty(self.world * 4, dtype=torch.int64, device="cuda")
        with torch.cuda.stream(self.stream):
            dist.all_gather_into_tensor(every, mine)
        self.stream.synchronize()
        every = every.view(self.world, 4).tolist()
        slot = max(1, max(sum(c) for c in every))
        send = torch.zeros(slot, dtype=torch.int64, device="cuda")
        recv = torch.empty(slot * self.world, dtype=torch.int64, device="cuda")
        if sum(n) > 0:
            lib.check(lib.collisions_corrections_pack_dev(self.ctx, C.c_void_p(send.data_ptr())))
        with torch.cuda.stream(self.stream):
            dist.all_gather_into_tensor(recv, send)
            parts, totals = [], []
            for k in range(4):  # list k of every rank, one after the other
                for r in range(self.world):
                    off = r * slot + sum(every[r][:k])
                    parts.append(recv[off:off + every[r][k]])
                totals.append(sum(every[r][k] for r in range(self.world)))
            keys = torch.cat(parts) if sum(totals) > 0 else torch.zeros(1, dtype=torch.int64, device="cuda")
        self.stream.synchronize()
        lib.check(lib.collisions_corrections_apply_dev(self.ctx, C.c_void_p(keys.data_ptr()), (C.c_int64 * 4)(*totals), self.counts))
        self._done()

    def step(self, dV0, dV1, d_energy, d_grad, d_step, dhat, bp, ccd, dmin=0.0, min_distance=0.0, psd=1, flags=0):
        """flags: IPCB_USE_AREA_WEIGHTING (1) | IPCB_SET_IMPROVED_MAX_APPROX (2); the latter needs the row-block form"""
        torch, lib, dist, ctx = self.torch, self.lib, self.dist, self.ctx
        self.improved = bool(flags & 2) and self.world > 1
        if self.improved and not self.row_block:
            raise ValueError("IMPROVED_MAX_APPROX over ranks needs the united set (row_block=True)")
        nV = self.mesh.num_vertices()
        p0, p1 = C.c_void_p(dV0.data_ptr()), C.c_void_p(dV1.data_ptr())
        pstep = C.c_void_p(d_step.data_ptr())

        def ccd_half(c):
            lib.check(lib.ccd_stepsize_dev(c, p0, p1, nV, min_distance, C.byref(ccd), pstep))
            self._done(c)

        job = None
        if self.lane is not None:  # the CCD half on the second context, beside everything below
            self.ev_start.record(self.stream)
            self.stream_b.wait_event(self.ev_start)  # the caller's inputs are ready when the first lane's stream gets here
            job = self.lane.submit(lambda: ccd_half(self.ctx_b))
        lib.check(lib.collisions_build_dev(ctx, p0, nV, dhat, dmin, flags, self.counts))
        self._done()
        if self.improved:
            self.exchange_correction_keys()
        if self.row_block:
            self.start_exchange()
        else:
            self.shard_counts = list(self.counts)
        if job is None:  # one context: the step size runs here, while the records travel
            ccd_half(ctx)
        if self.row_block:
            self.finish_exchange(dmin)
            lib.check(lib.ctx_set_collision_range(ctx, self.rank, self.world))
            lib.check(lib.ctx_set_row_block(ctx, *self.rows))
        lib.check(lib.barrier_energy_dev(ctx, p0, nV, C.byref(bp), C.c_void_p(d_energy.data_ptr())))
        self._done()
        lib.check(lib.barrier_gradient_dev(ctx, p0, nV, C.byref(bp), C.c_void_p(d_grad.data_ptr())))
        self._done()
        lib.check(lib.barrier_hessian_dev(ctx, p0, nV, C.byref(bp), psd, C.byref(self.nnz)))
        self._done()
        if job is not None:  # join the lanes: the first lane's stream continues after the step size has been written
            job.wait()
            self.ev_ccd.record(self.stream_b)
            self.stream.wait_event(self.ev_ccd)
        if self.world > 1:  # sum / sum / min all-reduces over NVLink (one coalesced group); the Hessian needs no collective
            with torch.cuda.stream(self.stream):
                dist.all_reduce(d_energy)
                dist.all_reduce(d_grad)
                dist.all_reduce(d_step, op=dist.ReduceOp.MIN)
        return self.nnz.value
