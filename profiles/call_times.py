"""Per-API-call device times over many steps of a bench workload (mean / min / max per call), L2 flushed between
steps like bench.py.   usage: python profiles/call_times.py c3 [steps]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import ipctk_b200  # noqa: E402


def main():
    wl = sys.argv[1]
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    api = ipctk_b200.library()
    lib, abi, scenes = api.lib, ipctk_b200._pkg._abi, ipctk_b200._pkg.scenes
    desc, spec, _ = bench.WORKLOADS[wl]
    V0, V1, E, F, P = bench.make_scene(scenes, spec)
    nV, dhat = V0.shape[0], P["dhat"]
    mesh = api.CollisionMesh(V0, E, F)
    ctx = mesh._ctx
    dV0 = torch.from_numpy(np.asfortranarray(V0).T.copy()).cuda()
    dV1 = torch.from_numpy(np.asfortranarray(V1).T.copy()).cuda()
    d_e = torch.zeros(1, dtype=torch.float64, device="cuda")
    d_g = torch.zeros(3 * nV, dtype=torch.float64, device="cuda")
    d_s = torch.zeros(1, dtype=torch.float64, device="cuda")
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    counts, nnz = (C.c_int64 * 4)(), C.c_int64()
    bp, ccd = abi.BarrierParams(dhat, 1.0, 0), abi.CcdParams(0, 0.0, 0, 0.0)
    p0, p1 = C.c_void_p(dV0.data_ptr()), C.c_void_p(dV1.data_ptr())
    stream = torch.cuda.ExternalStream(lib.ctx_stream(ctx))
    calls = [
        ("collisions_build", lambda: lib.collisions_build_dev(ctx, p0, nV, dhat, 0.0, 0, counts)),
        ("energy", lambda: lib.barrier_energy_dev(ctx, p0, nV, C.byref(bp), C.c_void_p(d_e.data_ptr()))),
        ("gradient", lambda: lib.barrier_gradient_dev(ctx, p0, nV, C.byref(bp), C.c_void_p(d_g.data_ptr()))),
        ("hessian", lambda: lib.barrier_hessian_dev(ctx, p0, nV, C.byref(bp), 1, C.byref(nnz))),
        ("swept_candidates", lambda: lib.candidates_build_swept_dev(ctx, p0, p1, nV, 0.0, counts)),
        ("ccd_from_candidates", lambda: lib.ccd_stepsize_from_candidates_dev(ctx, p0, p1, nV, 0.0, C.byref(ccd), C.c_void_p(d_s.data_ptr()))),
    ]
    T = {n: [] for n, _ in calls}
    T["step"] = []
    for it in range(steps + 3):
        flush.fill_(1)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(calls) + 1)]
        ev[0].record(stream)
        for k, (name, fn) in enumerate(calls):
            lib.check(fn())
            ev[k + 1].record(stream)
        ev[-1].synchronize()
        if it >= 3:
            for k, (name, _) in enumerate(calls):
                T[name].append(ev[k].elapsed_time(ev[k + 1]))
            T["step"].append(ev[0].elapsed_time(ev[-1]))
    for name, v in T.items():
        v = np.array(v)
        print("%-22s mean %7.3f  min %7.3f  p50 %7.3f  max %7.3f" % (name, v.mean(), v.min(), np.median(v), v.max()))


if __name__ == "__main__":
    main()
