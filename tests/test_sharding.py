"""Multi-GPU host logic on CPU: two gloo ranks, each working on its shard of the candidates through
the oracle library, must reproduce the single-rank energy / gradient / step size after the all-reduces, and the sum
of the rank Hessians must equal the single-rank Hessian (same pattern, values to 1e-10).  The GPU variant runs the
same class on NCCL with the device-side Morton-range shard (`-m gpu`, needs 2 GPUs)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, backend, queue):
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
        out = sharded.ShardedContactStep(api, mesh, rank, world, dist=dist, device=device, native=backend == "nccl").step(
            V0, V1, P["dhat"])
        H = out["hessian_local"].toarray()
        t = torch.from_numpy(H)
        if device is not None:
            t = t.to(device)
        dist.all_reduce(t)  # test only: the product keeps the Hessian per rank
        # single-rank reference on the same library
        mesh1 = api.CollisionMesh(V0, E, F, **({"device": rank} if backend == "nccl" else {}))
        one = sharded.ShardedContactStep(api, mesh1, 0, 1, native=backend == "nccl").step(V0, V1, P["dhat"])
        H1 = one["hessian_local"].toarray()
        Hs = t.cpu().numpy()
        res = dict(
            rank=rank, shard_collisions=out["collisions"], all_collisions=one["collisions"],
            energy=abs(out["energy"] - one["energy"]) / abs(one["energy"]),
            grad=np.linalg.norm(out["gradient"] - one["gradient"]) / np.linalg.norm(one["gradient"]),
            step=(out["step"], one["step"]),
            hess=np.linalg.norm(Hs - H1) / np.linalg.norm(H1), pattern=bool(np.array_equal(Hs != 0, H1 != 0)))
        dist.destroy_process_group()
        queue.put(res)
    except Exception as e:  # pragma: no cover
        import traceback

        queue.put(dict(rank=rank, error=traceback.format_exc() + repr(e)))


def _run(backend, world=2):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, backend, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(60)
    for r in results:
        assert "error" not in r, r.get("error")
    return sorted(results, key=lambda r: r["rank"])


def _check(results):
    total = np.sum([r["shard_collisions"] for r in results], axis=0)
    # the shards partition the candidates: FV / EE collisions are disjoint, VV / EV ones may be derived on both ranks
    assert total[2] == results[0]["all_collisions"][2] and total[3] == results[0]["all_collisions"][3]
    assert total[0] >= results[0]["all_collisions"][0] and total[1] >= results[0]["all_collisions"][1]
    assert all(sum(r["shard_collisions"]) > 0 for r in results)
    for r in results:
        assert r["energy"] <= 1e-12 and r["grad"] <= 1e-12
        assert r["step"][0] == pytest.approx(r["step"][1], rel=1e-3, abs=1e-6)
        assert r["hess"] <= 1e-10 and r["pattern"]


def test_two_rank_contact_step_gloo(oracle):  # the fixture builds the oracle before the ranks race for it
    _check(_run("gloo"))


@pytest.mark.gpu
def test_two_rank_contact_step_nccl(cuda):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _check(_run("nccl"))
