"""Timing sweep of the step-size search alone (swept broad phase excluded) on a bench workload.
usage: python profiles/ccd_sweep.py c3 "IPCB_TI_SAMPLE=65536" "IPCB_TI_SAMPLE=8192,IPCB_TI_BUDGET=64" ...
Each argument is one configuration (comma-separated environment variables read by the library per call)."""
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
    configs = sys.argv[2:] or [""]
    api = ipctk_b200.library()
    lib, abi, scenes = api.lib, ipctk_b200._pkg._abi, ipctk_b200._pkg.scenes
    desc, spec, _ = bench.WORKLOADS[wl]
    V0, V1, E, F, P = bench.make_scene(scenes, spec)
    nV = V0.shape[0]
    mesh = api.CollisionMesh(V0, E, F)
    ctx = mesh._ctx
    dV0 = torch.from_numpy(np.asfortranarray(V0).T.copy()).cuda()
    dV1 = torch.from_numpy(np.asfortranarray(V1).T.copy()).cuda()
    d_step = torch.zeros(1, dtype=torch.float64, device="cuda")
    counts = (C.c_int64 * 4)()
    ccd = abi.CcdParams(0, 0.0, 0, 0.0)
    p0, p1 = C.c_void_p(dV0.data_ptr()), C.c_void_p(dV1.data_ptr())
    lib.check(lib.candidates_build_swept_dev(ctx, p0, p1, nV, 0.0, counts))
    print("candidates", list(counts))
    stream = torch.cuda.ExternalStream(lib.ctx_stream(ctx))
    for cfg in configs:
        env = dict(kv.split("=") for kv in cfg.split(",") if kv)
        os.environ.update(env)
        ts = []
        for it in range(25):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record(stream)
            lib.check(lib.ccd_stepsize_from_candidates_dev(ctx, p0, p1, nV, 0.0, C.byref(ccd), C.c_void_p(d_step.data_ptr())))
            b.record(stream)
            b.synchronize()
            if it >= 5:
                ts.append(a.elapsed_time(b))
        for k in env:
            del os.environ[k]
        ts = np.array(ts)
        print("%-50s mean %.3f  min %.3f  max %.3f  step %.6f" % (cfg or "default", ts.mean(), ts.min(), ts.max(), float(d_step.item())))


if __name__ == "__main__":
    main()
