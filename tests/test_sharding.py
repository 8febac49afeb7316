"""Multi-GPU host logic on CPU: two gloo ranks, each working on its shard of the candidates through
the oracle library, must reproduce the single-rank energy / gradient / step size after the all-reduces.
row_block mode: the ranks' collision records are all-gathered and merged (every rank then holds exactly the
single-rank set), energy / gradient run on collision ranges and every rank assembles only its row block of the
Hessian: the blocks must tile the single-rank matrix (pattern identical, values to 1e-10).  additive mode: the sum
of the rank Hessians must equal the single-rank Hessian.  The GPU variants run the same classes on NCCL with the
device-side Morton-range shard (`-m gpu`, need 2 GPUs), and the device-resident step `bench.py` uses."""
import os
import socket
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, backend, row_block, queue, improved=False):
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        import torch
        import torch.distributed as dist

        import ipctk_b200

        scenes = ipctk_b200._pkg.scenes
        sharded = __import__("importlib").import_module("ipc_toolkit_b200.sharded")
        V0, V1, E, F, P = scenes.cloth_stack(3, 14)
        device = None
        if backend == "nccl":
            torch.cuda.set_device(rank)
            device = torch.device("cuda", rank)
            dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
            api = ipctk_b200.library()
            mesh = api.CollisionMesh(V0, E, F, device=rank)
        else:
            dist.init_process_group("gloo", rank=rank, world_size=world)
            import pyoracle

            api = pyoracle.load()
            api.set_num_threads(2)
            mesh = api.CollisionMesh(V0, E, F)
        out = sharded.ShardedContactStep(api, mesh, rank, world, dist=dist, device=device, native=backend == "nccl",
                                         row_block=row_block).step(V0, V1, P["dhat"], improved_max_approx=improved)
        H = np.ascontiguousarray(out["hessian_local"].toarray())
        Hmine = H.copy()
        t = torch.from_numpy(H)
        if device is not None:
            t = t.to(device)
        dist.all_reduce(t)  # test only: the product keeps the Hessian per rank
        # single-rank reference on the same library
        mesh1 = api.CollisionMesh(V0, E, F, **({"device": rank} if backend == "nccl" else {}))
        one = sharded.ShardedContactStep(api, mesh1, 0, 1, native=backend == "nccl").step(V0, V1, P["dhat"], improved_max_approx=improved)
        H1 = one["hessian_local"].toarray()
        Hs = t.cpu().numpy()
        tiles = None
        if row_block:  # the rank's matrix is exactly the single-rank matrix restricted to its columns (== rows)
            lo, hi = out["rows"]
            mask = np.zeros(H1.shape[1], bool)
            mask[3 * lo:3 * hi] = True
            inside = H1[:, mask]
            tiles = dict(rows=(lo, hi), outside_empty=not Hmine[:, ~mask].any(),
                         pattern=bool(np.array_equal(Hmine[:, mask] != 0, inside != 0)),
                         err=float(np.linalg.norm(Hmine[:, mask] - inside) / max(np.linalg.norm(inside), 1e-300)),
                         same_set=out["collisions"] == one["collisions"])
        res = dict(
            rank=rank, shard_collisions=out["shard_collisions"], all_collisions=one["collisions"], tiles=tiles,
            energy=abs(out["energy"] - one["energy"]) / abs(one["energy"]),
            grad=np.linalg.norm(out["gradient"] - one["gradient"]) / np.linalg.norm(one["gradient"]),
            step=(out["step"], one["step"]),
            hess=np.linalg.norm(Hs - H1) / np.linalg.norm(H1), pattern=bool(np.array_equal(Hs != 0, H1 != 0)))
        dist.destroy_process_group()
        queue.put(res)
    except Exception as e:  # pragma: no cover
        import traceback

        queue.put(dict(rank=rank, error=traceback.format_exc() + repr(e)))


def _run(backend, world=2, row_block=True, improved=False):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, backend, row_block, q, improved)) for r in range(world)]
    for p in procs:
        p.start()
    return _collect(q, procs)


def _collect(q, procs, timeout=240):
    """results of all ranks; a rank that failed (its peers then wait in a collective forever) fails the test at once"""
    import queue as queue_mod

    results, error = [], None
    try:
        for _ in procs:
            r = q.get(timeout=timeout)
            if "error" in r:
                error = r["error"]
                break
            results.append(r)
    except queue_mod.Empty:
        error = "timeout waiting for the ranks"
    for p in procs:
        if error is None:
            p.join(60)
        if p.is_alive():
            p.kill()
    assert error is None, error
    return sorted(results, key=lambda r: r["rank"])


def _step_tolerance(make_scene, steps):
    """derived tolerance between two valid Tight-Inclusion step sizes on the scene (tests/ccd_tolerance.py)"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ipctk_b200
    import pyoracle
    from ccd_tolerance import StepTolerance

    V0, V1, E, F, P = make_scene(ipctk_b200._pkg.scenes)
    return StepTolerance(pyoracle.load(), V0, V1, E, F).tolerance(min(steps))[0]


def _check(results, improved=False):
    tol = _step_tolerance(lambda s: s.cloth_stack(3, 14), [x for r in results for x in r["step"]])
    total = np.sum([r["shard_collisions"] for r in results], axis=0)
    if not improved:  # (the correction records of IMPROVED_MAX_APPROX cancel and coincide across ranks)
        # the shards partition the candidates: FV / EE collisions are disjoint, VV / EV ones may be derived on both ranks
        assert total[2] == results[0]["all_collisions"][2] and total[3] == results[0]["all_collisions"][3]
        assert total[0] >= results[0]["all_collisions"][0] and total[1] >= results[0]["all_collisions"][1]
    assert all(sum(r["shard_collisions"]) > 0 for r in results)
    for r in results:
        assert r["energy"] <= 1e-12 and r["grad"] <= 1e-12
        assert abs(r["step"][0] - r["step"][1]) <= tol, (r["step"], tol)
        assert r["hess"] <= 1e-10 and r["pattern"]
        if r["tiles"] is not None:
            t = r["tiles"]
            assert t["same_set"] and t["outside_empty"] and t["pattern"] and t["err"] <= 1e-10, t
    if results[0]["tiles"] is not None:  # the row blocks partition the vertices
        rows = [r["tiles"]["rows"] for r in results]
        assert rows[0][0] == 0 and all(rows[k][1] == rows[k + 1][0] for k in range(len(rows) - 1))
        assert all(hi > lo for lo, hi in rows)


@pytest.mark.parametrize("row_block", [True, False], ids=["row_block", "additive"])
def test_two_rank_contact_step_gloo(oracle, row_block):  # the fixture builds the oracle before the ranks race for it
    _check(_run("gloo", row_block=row_block))


def test_two_rank_improved_max_approx_gloo(oracle):
    """CollisionSetType::IMPROVED_MAX_APPROX over two ranks (deferred build, exchange of the sub-element pairs, per-rank slices
    of the corrections: include/ipcb200.h, IPCB_DEFER_CORRECTIONS) equals the single-rank step with the same set type"""
    _check(_run("gloo", row_block=True, improved=True), improved=True)


@pytest.mark.gpu
def test_two_rank_improved_max_approx_nccl_host_buffers(cuda):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _check(_run("nccl", row_block=True, improved=True), improved=True)


@pytest.mark.gpu
@pytest.mark.parametrize("row_block", [True, False], ids=["row_block", "additive"])
def test_two_rank_contact_step_nccl(cuda, row_block):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _check(_run("nccl", row_block=row_block))


def _device_worker(rank, world, port, queue, flags=0):
    """the device-resident sharded step of bench.py (DeviceShardedStep) against the single-context step on the same GPU"""
    try:
        import ctypes as C

        sys.path.insert(0, ROOT)
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        import scipy.sparse as sp
        import torch
        import torch.distributed as dist

        import ipctk_b200

        scenes, abi = ipctk_b200._pkg.scenes, ipctk_b200._pkg._abi
        sharded = __import__("importlib").import_module("ipc_toolkit_b200.sharded")
        torch.cuda.set_device(rank)
        device = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
        api = ipctk_b200.library()
        lib = api.lib
        V0, V1, E, F, P = scenes.cloth_stack(4, 40)
        nV = V0.shape[0]
        dV0 = torch.from_numpy(np.asfortranarray(V0).T.copy()).cuda()
        dV1 = torch.from_numpy(np.asfortranarray(V1).T.copy()).cuda()
        bp, ccd = abi.BarrierParams(P["dhat"], 1.0, 0), abi.CcdParams(0, 0.0, 0, 0.0)

        keep, steppers = [], []  # contexts (and their streams) must outlive every tensor used on them

        def run(r, w, d, two_lanes):
            mesh = api.CollisionMesh(V0, E, F, device=rank)
            ccd_mesh = api.CollisionMesh(V0, E, F, device=rank) if two_lanes else None
            keep.extend([mesh, ccd_mesh])
            stream = torch.cuda.ExternalStream(lib.ctx_stream(mesh._ctx), device=device)
            st = sharded.DeviceShardedStep(api, mesh, r, w, d, torch, stream, ccd_mesh=ccd_mesh)
            steppers.append(st)
            e, g, s = (torch.zeros(n, dtype=torch.float64, device="cuda") for n in (1, 3 * nV, 1))
            for _ in range(2):  # twice: buffers are reused across steps
                nnz = st.step(dV0, dV1, e, g, s, P["dhat"], bp, ccd, flags=flags)
            torch.cuda.synchronize()
            outer, inner, vals = np.zeros(3 * nV + 1, np.int32), np.zeros(nnz, np.int32), np.zeros(nnz)
            lib.check(lib.barrier_hessian_fetch(mesh._ctx, outer.ctypes.data_as(C.c_void_p), inner.ctypes.data_as(C.c_void_p),
                                                vals.ctypes.data_as(C.c_void_p)))
            H = sp.csc_matrix((vals, inner, outer), shape=(3 * nV, 3 * nV))
            return dict(e=float(e.item()), g=g.cpu().numpy(), s=float(s.item()), H=H, rows=st.rows, counts=list(st.counts),
                        shard=list(st.shard_counts))

        out = run(rank, world, dist, True)  # the bench's configuration: CCD half on a second context
        one = run(0, 1, None, False)
        lo, hi = 3 * out["rows"][0], 3 * out["rows"][1]
        A, B = out["H"][:, lo:hi], one["H"][:, lo:hi]
        res = dict(rank=rank, rows=out["rows"], same_set=out["counts"] == one["counts"], shard=out["shard"],
                   energy=abs(out["e"] - one["e"]) / abs(one["e"]), grad=float(np.linalg.norm(out["g"] - one["g"]) / np.linalg.norm(one["g"])),
                   step=(out["s"], one["s"]), outside=int(out["H"][:, :lo].nnz + out["H"][:, hi:].nnz),
                   pattern=bool(np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)),
                   hess=float(np.linalg.norm(A.data - B.data) / np.linalg.norm(B.data)) if A.nnz == B.nnz else 1.0, nnz=int(A.nnz),
                   nnz_all=int(one["H"].nnz))
        torch.cuda.synchronize()
        # orderly teardown: tensors and torch's cached blocks first, then NCCL, then the contexts with their streams
        for st in steppers:
            st.release()
        del steppers, dV0, dV1
        import gc

        gc.collect()
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
        dist.destroy_process_group()
        for m in keep:
            if m is not None:
                m.close()
        queue.put(res)
    except Exception as e:  # pragma: no cover
        import traceback

        queue.put(dict(rank=rank, error=traceback.format_exc() + repr(e)))


@pytest.mark.gpu
def test_device_sharded_step_nccl(cuda):
    import torch
    import torch.multiprocessing as mp

    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_device_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = _collect(q, procs)
    tol = _step_tolerance(lambda s: s.cloth_stack(4, 40), [x for r in results for x in r["step"]])
    for r in results:
        assert r["same_set"] and r["outside"] == 0 and r["pattern"] and r["hess"] <= 1e-10, r
        assert r["energy"] <= 1e-12 and r["grad"] <= 1e-12
        assert abs(r["step"][0] - r["step"][1]) <= tol, (r["step"], tol)
        assert sum(r["shard"]) > 0
    rows = [r["rows"] for r in results]
    assert rows[0][0] == 0 and all(rows[k][1] == rows[k + 1][0] for k in range(world - 1))
    assert sum(r["nnz"] for r in results) == results[0]["nnz_all"]  # the row blocks tile the matrix


@pytest.mark.gpu
def test_device_sharded_step_improved_max_approx_nccl(cuda):
    """CollisionSetType::IMPROVED_MAX_APPROX over NCCL ranks: the sub-element keys are exchanged before the corrections
    (DeviceShardedStep.exchange_correction_keys), and the step equals the single-context step with the same set type"""
    import torch
    import torch.multiprocessing as mp

    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_device_worker, args=(r, world, port, q, 2)) for r in range(world)]
    for p in procs:
        p.start()
    results = _collect(q, procs)
    tol = _step_tolerance(lambda s: s.cloth_stack(4, 40), [x for r in results for x in r["step"]])
    for r in results:
        assert r["same_set"] and r["outside"] == 0 and r["pattern"] and r["hess"] <= 1e-10, r
        assert r["energy"] <= 1e-12 and r["grad"] <= 1e-12
        assert abs(r["step"][0] - r["step"][1]) <= tol, (r["step"], tol)
    assert sum(r["nnz"] for r in results) == results[0]["nnz_all"]


def _empty_worker(rank, world, port, queue):
    """no rank finds a collision: the exchange, the merge, the balanced row blocks and the potential must cope with empty
    sets (energy 0, zero gradient, empty Hessian) while the step size still comes from the swept candidates"""
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        import torch.distributed as dist

        import ipctk_b200

        scenes = ipctk_b200._pkg.scenes
        sharded = __import__("importlib").import_module("ipc_toolkit_b200.sharded")
        V0, V1, E, F, P = scenes.cloth_stack(2, 6, gap=50.0)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import pyoracle

        api = pyoracle.load()
        api.set_num_threads(2)
        out = sharded.ShardedContactStep(api, api.CollisionMesh(V0, E, F), rank, world, dist=dist, native=False).step(V0, V1, P["dhat"])
        one = sharded.ShardedContactStep(api, api.CollisionMesh(V0, E, F), 0, 1, native=False).step(V0, V1, P["dhat"])
        res = dict(rank=rank, collisions=out["collisions"], energy=out["energy"], grad=float(np.abs(out["gradient"]).max()),
                   nnz=int(out["hessian_local"].nnz), rows=out["rows"], step=(out["step"], one["step"]))
        dist.destroy_process_group()
        queue.put(res)
    except Exception as e:  # pragma: no cover
        import traceback

        queue.put(dict(rank=rank, error=traceback.format_exc() + repr(e)))


def test_two_rank_step_without_collisions_gloo(oracle):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_empty_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = _collect(q, procs)
    tol = _step_tolerance(lambda s: s.cloth_stack(2, 6, gap=50.0), [x for r in results for x in r["step"]])
    for r in results:
        assert r["collisions"] == [0, 0, 0, 0] and r["energy"] == 0.0 and r["grad"] == 0.0 and r["nnz"] == 0
        assert 0 < r["step"][0] < 1 and abs(r["step"][0] - r["step"][1]) <= tol, (r["step"], tol)
    rows = [r["rows"] for r in results]
    assert rows[0][0] == 0 and rows[0][1] == rows[1][0]
