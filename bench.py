"""bench.py — ms per contact step (NormalCollisions::build + barrier E / grad / Hessian(CLAMP) +
compute_collision_free_stepsize) on the BASELINE.json workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c2|c1|c5]

One JSON line on stdout (rank 0).  `value` = whole-job ms per step with inputs resident in HBM,
timed with CUDA events on the library's own stream, max over ranks; `e2e` = the same step through the
host-buffer C ABI (pinned host inputs, host results); `roofline` = the dominant kernel against the
measured HBM copy bandwidth; `cpu_baseline` = the CPU restatement on a bounded sample of the workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ms per contact step (build+barrier E/grad/Hess+CCD)"

WORKLOADS = {
    # name: (description, generator kwargs, bounded CPU sample kwargs, sample scale = full tris / sample tris)
    "c1": ("C1 64x64 cloth over a 32x32 UV sphere (~10K tris), dhat=1e-3", dict(kind="sphere", n=64, res=32, drape=False), None),
    "c2": ("C2 256x256 cloth draped on a 50K-tri sphere (~180K tris), dhat=1e-3", dict(kind="sphere", n=256, res=160, drape=True), None),
    "c3": ("C3 8 stacked 250x250 cloth layers (1.0M tris), gap 0.5*dhat, dense edge-edge contact", dict(kind="stack", layers=8, n=250, gap=0.5),
           dict(kind="stack", layers=3, n=100, gap=0.5, h=1.0 / 250)),
    "c5": ("C5 16-layer compressed stack (2.0M tris), gap 0.2*dhat, PSD=CLAMP", dict(kind="stack", layers=16, n=250, gap=0.2),
           dict(kind="stack", layers=3, n=100, gap=0.2, h=1.0 / 250)),
    # BASELINE config 4: broad phase + CCD stress — only the collision-free step size is the step
    "c4": ("C4 20 perturbed 500x500 sheets (10.0M tris), vertices displaced U(-1,1)*2h: swept broad phase + CCD step size only",
           dict(kind="sheets", layers=20, n=500, spacing=2.0, disp=2.0, ccd_only=True), None),
    "c4s": ("C4 at 1/10 size: 8 perturbed 250x250 sheets (1.0M tris), vertices displaced U(-1,1)*2h: swept broad phase + CCD step size only",
            dict(kind="sheets", layers=8, n=250, spacing=2.0, disp=2.0, ccd_only=True), None),
}


def make_scene(scenes, spec):
    if spec["kind"] == "sphere":
        return scenes.cloth_on_sphere(spec["n"], spec["res"], drape=spec["drape"])
    if spec["kind"] == "sheets":
        return scenes.perturbed_sheets(spec["layers"], spec["n"], spacing=spec["spacing"], disp=spec["disp"])
    return scenes.cloth_stack(spec["layers"], spec["n"], gap=spec["gap"], h=spec.get("h"))


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """samples nvidia-smi clocks and throttle reasons while the timed region runs"""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(s[2 + k].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.samples[0][1]), "reasons": reasons,
                "samples": len(self.samples)}


def ncu_traffic(workload):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the committed `ncu --set full` capture of
    this workload (profiles/r2_ncu_full_<workload>_raw.csv), by kernel name; None when there is no capture"""
    import csv

    path = os.path.join(ROOT, "profiles", "r2_ncu_full_%s_raw.csv" % workload)
    if not os.path.exists(path):
        return None
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    kn, rd, wr = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    out = {}
    for r in rows[2:]:
        name = r[kn].replace("ipcb::", "").replace("void ", "").split("(")[0]
        name = name.replace("(int)", "").replace(", ", ",")
        out.setdefault(name, 0.0)
        out[name] += float(r[rd].replace(",", "")) * scale.get(units[rd], 1.0) + float(r[wr].replace(",", "")) * scale.get(units[wr], 1.0)
    return out


def cpu_step(api, mesh, V0, V1, dhat, ccd_only=False):
    """one contact step through the oracle's (reference-equivalent) CPU path"""
    if ccd_only:
        import scipy.sparse as sp

        step = api.compute_collision_free_stepsize(mesh, V0, V1)
        return 0.0, None, sp.csc_matrix((1, 1)), step, [0, 0, 0, 0]
    c = api.NormalCollisions()
    c.build(mesh, V0, dhat)
    B = api.BarrierPotential(dhat, 1.0)
    e = B(c, mesh, V0)
    g = B.gradient(c, mesh, V0)
    H = B.hessian(c, mesh, V0, api.PSDProjectionMethod.CLAMP)
    step = api.compute_collision_free_stepsize(mesh, V0, V1)
    return e, g, H, step, c.counts()


def run_reference(args, desc, full_spec, sample_spec):
    """--impl reference: the reference's CPU path for this step — the restatement in oracle/ (the upstream library cannot
    be built here: no Eigen / TBB / Tight-Inclusion in the image) — on the FULL workload of the GPU arm, with
    every host thread.  One step of C3 takes tens of seconds on the host, so the number of steps is bounded by a time
    budget (IPCB_REFERENCE_BUDGET_S, default 240 s): the line reports the steps that really ran, never a scaled value."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    import importlib.util

    spec = importlib.util.spec_from_file_location("ipcb_scenes", os.path.join(ROOT, "ipc-toolkit_b200", "scenes.py"))
    scenes = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(scenes)
    api = pyoracle.load(fast=True)
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    api.set_num_threads(cores)  # torchrun exports OMP_NUM_THREADS=1: the baseline uses every core at every N
    V0, V1, E, F, P = make_scene(scenes, full_spec)
    mesh = api.CollisionMesh(V0, E, F)
    budget = float(os.environ.get("IPCB_REFERENCE_BUDGET_S", "240"))
    t_start = time.perf_counter()
    times, warm = [], 0
    out = None
    n_warm = min(args.warmup, 1)  # one warm-up step (allocations, page-in); the rest of the budget goes to timed steps
    while len(times) < args.steps:
        t = time.perf_counter()
        out = cpu_step(api, mesh, V0, V1, P["dhat"], full_spec.get("ccd_only", False))
        dt = time.perf_counter() - t
        if warm < n_warm:
            warm += 1
        else:
            times.append(dt * 1e3)
        if len(times) >= 2 and (time.perf_counter() - t_start) + dt > budget:
            break
    ms = float(np.mean(times))
    sample = ("the full workload (%d triangles); %d timed steps after %d warm-up steps inside a %.0f s budget (%d / %d requested)"
              % (F.shape[0], len(times), warm, budget, args.steps, args.warmup))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": ms, "unit": "ms", "n_gpus": args.gpus, "steps": len(times), "warmup": warm,
        "steps_requested": args.steps, "warmup_requested": args.warmup,
        "ms_per_step": ms, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": make_config(desc, full_spec, args.gpus, args.additive_hessian),
        "cpu_baseline": {"value": ms, "unit": "ms", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "counts": {"collisions": out[4], "hessian_nnz": int(out[2].nnz), "step": out[3], "energy": out[0]},
    }))


def make_config(desc, full_spec, world, additive=False):
    """the `config` object — identical for both arms (the reference arm times the SAME workload on the host cores)"""
    nV, nE, nF = scene_sizes(full_spec)
    return {"workload": desc, "triangles": nF, "vertices": nV, "edges": nE, "dhat": 1e-3 if full_spec["kind"] != "sheets" else None,
            "psd": "CLAMP", "ccd": "TightInclusion",
            "collision_set": "IPC", "l2": "GPU arm: flushed between timed steps (256 MB write); CPU arm: inputs larger than the caches",
            "parallelism": ("single GPU" if world == 1 else
                            "candidate shards by Morton range of query leaves; " +
                            ("Hessian as additive rank contributions" if additive else
                             "collision records all-gathered + merged, energy/gradient by collision range, "
                             "Hessian by balanced row block (no collective)"))}


def scene_sizes(spec):
    """(vertices, edges, triangles) of a generator spec without building it"""
    if spec["kind"] == "sphere":
        n, res = spec["n"], spec["res"]
        nV = (n + 1) ** 2 + res * (res - 1) + 2
        nF = 2 * n * n + 2 * res * (res - 1)
        nE = (3 * n * n + 2 * n) + 3 * res * (res - 1)
        return nV, nE, nF
    L, n = spec["layers"], spec["n"]
    return L * (n + 1) ** 2, L * (3 * n * n + 2 * n), L * 2 * n * n


def make_tris(spec):
    if spec["kind"] == "sphere":
        return 2 * spec["n"] ** 2 + 2 * spec["res"] * (spec["res"] - 1)
    return spec["layers"] * 2 * spec["n"] ** 2


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--additive-hessian", action="store_true",
                    help="N > 1: keep every rank's potential on its own collision shard (Hessian = additive contribution) "
                         "instead of the all-gathered set with row-block Hessians")
    ap.add_argument("--single-context", action="store_true",
                    help="A/B: run the five calls of the step one after the other on ONE context (no second lane for the CCD half)")
    ap.add_argument("--broad", default="lbvh", choices=["lbvh", "sap"], help="broad-phase method of the CUDA library (A/B: LBVH or sweep-and-prune)")
    ap.add_argument("--ncu-step", action="store_true",
                    help="profiling aid: warm up, then run ONE device step between cudaProfilerStart/Stop and exit "
                         "(use with ncu --profile-from-start off); prints no bench line")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    desc, full_spec, sample_spec = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, desc, full_spec, sample_spec)

    import torch
    import ipctk_b200

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    json_fd = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("NCCL_DEBUG", "WARN")
        # NCCL prints its version banner on stdout: point fd 1 at stderr for the run and keep the real stdout for the
        # ONE JSON line
        sys.stdout.flush()
        json_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    api = ipctk_b200.library()
    lib = api.lib
    scenes = ipctk_b200._pkg.scenes
    abi = ipctk_b200._pkg._abi
    V0, V1, E, F, P = make_scene(scenes, full_spec)
    dhat = P["dhat"]
    nV = V0.shape[0]
    assert (nV, E.shape[0], F.shape[0]) == scene_sizes(full_spec)
    mesh = api.CollisionMesh(V0, E, F, device=local)
    ctx = mesh._ctx
    # second context on the same mesh: the CCD half of the step runs on its own streams beside the potential half
    ccd_only = bool(full_spec.get("ccd_only"))
    ccd_mesh = None if (args.single_context or ccd_only) else api.CollisionMesh(V0, E, F, device=local)
    for m in (mesh, ccd_mesh):
        if m is not None:
            m.set_broad_phase_method(args.broad)
    contexts = [ctx] + ([ccd_mesh._ctx] if ccd_mesh is not None else [])
    lib.check(lib.ctx_set_shard(ctx, rank, world))
    stream = torch.cuda.ExternalStream(lib.ctx_stream(ctx), device=torch.device("cuda", local))

    # ---- device-resident inputs (column-major N x 3 like Eigen)
    dV0 = torch.from_numpy(np.asfortranarray(V0).T.copy()).cuda()  # 3 x N contiguous == N x 3 column-major
    dV1 = torch.from_numpy(np.asfortranarray(V1).T.copy()).cuda()
    d_energy = torch.zeros(1, dtype=torch.float64, device="cuda")
    d_grad = torch.zeros(3 * nV, dtype=torch.float64, device="cuda")
    d_step = torch.zeros(1, dtype=torch.float64, device="cuda")
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    bp = abi.BarrierParams(dhat, 1.0, 0)
    ccd = abi.CcdParams(0, 0.0, 0, 0.0)
    counts = (C.c_int64 * 4)()
    nnz = C.c_int64()
    info = {}
    stage_acc = {}

    def collect_stages(c):
        names = (C.c_char_p * 64)()
        ms = (C.c_float * 64)()
        n = lib.ctx_stage_times(c, 64, names, ms)
        for i in range(n):
            stage_acc.setdefault(names[i].decode(), []).append(ms[i])

    def launch_count():
        total = 0
        for c in contexts:
            n = C.c_int64()
            lib.ctx_launch_count(c, C.byref(n))
            total += n.value
        return total

    # the step itself lives in the package (ipc-toolkit_b200/sharded.py) so that the tests exercise the same code:
    # N = 1: the five library calls (the CCD half on the second context); N > 1: sharded broad phase, all-gather + merge of
    # the collision records, energy / gradient by collision range, Hessian by balanced row block, all-reduces
    sharded = __import__("importlib").import_module("ipc_toolkit_b200.sharded")
    stepper = sharded.DeviceShardedStep(api, mesh, rank, world, dist, torch, stream, row_block=not args.additive_hessian, ccd_mesh=ccd_mesh)

    def ccd_only_step(record=False):
        lib.check(lib.ccd_stepsize_dev(ctx, C.c_void_p(dV0.data_ptr()), C.c_void_p(dV1.data_ptr()), nV, 0.0, C.byref(ccd), C.c_void_p(d_step.data_ptr())))
        if record:
            collect_stages(ctx)
        info["nnz"], info["collisions"], info["shard_collisions"], info["rows"] = 0, [0, 0, 0, 0], [0, 0, 0, 0], [0, nV]

    def device_step(record=False):
        if ccd_only:
            return ccd_only_step(record)
        stepper.after = collect_stages if record else None
        info["nnz"] = stepper.step(dV0, dV1, d_energy, d_grad, d_step, dhat, bp, ccd)
        info["collisions"] = list(stepper.counts)
        info["shard_collisions"] = list(stepper.shard_counts)
        info["rows"] = list(stepper.rows)
        nnz.value = info["nnz"]

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        total = 0.0
        for _ in range(steps):
            flush.fill_(1)  # evict L2 between timed iterations
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            fn()
            b.record(stream)
            b.synchronize()
            total += a.elapsed_time(b)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t = torch.tensor([total], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)  # max over ranks
        return float(t.item()) / steps

    if args.ncu_step:
        for _ in range(args.warmup):
            device_step()
        torch.cuda.synchronize()
        flush.fill_(1)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        device_step()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        stepper.release()
        return

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = launch_count()
    ms_dev = timed(device_step, args.steps, args.warmup)
    launches = (launch_count() - l0) // (args.steps + args.warmup) * args.steps
    for c in contexts:
        lib.ctx_enable_stage_timing(c, 1)
    device_step(record=True)  # one extra, untimed pass to read the per-stage / per-kernel device times
    for c in contexts:
        lib.ctx_enable_stage_timing(c, 0)
    lib.check(lib.candidates_build_swept_dev(ctx, C.c_void_p(dV0.data_ptr()), C.c_void_p(dV1.data_ptr()), nV, 0.0, counts))
    info["ccd_candidates"] = list(counts)
    if world == 1 and not ccd_only:
        hv = np.asfortranarray(V0)
        lib.check(lib.candidates_build_static(ctx, hv.ctypes.data_as(C.c_void_p), nV, 0.5 * dhat, counts))
        info["static_candidates"] = list(counts)

    # ---- the line-search inner loop of the reference's solver example (python/examples/solver.py:
    # 95-116) from RESIDENT candidates: swept candidates with inflation dhat built once, then per line-search point
    # NormalCollisions::build(candidates, mesh, X, dhat) + the barrier energy — no broad phase inside the loop
    line_search = None
    if world == 1 and not ccd_only:
        p0, p1 = C.c_void_p(dV0.data_ptr()), C.c_void_p(dV1.data_ptr())
        lib.check(lib.candidates_build_swept_dev(ctx, p0, p1, nV, dhat, counts))
        ls_cand = list(counts)
        lib.check(lib.ccd_stepsize_from_candidates_dev(ctx, p0, p1, nV, 0.0, C.byref(ccd), C.c_void_p(d_step.data_ptr())))
        torch.cuda.synchronize()
        alpha = float(d_step.item())
        dX = (dV0 + 0.5 * alpha * (dV1 - dV0)).contiguous()
        ls_counts = (C.c_int64 * 4)()

        def rebuild():
            lib.check(lib.collisions_build_from_candidates_dev(ctx, C.c_void_p(dX.data_ptr()), nV, dhat, 0.0, 0, ls_counts))
            lib.check(lib.barrier_energy_dev(ctx, C.c_void_p(dX.data_ptr()), nV, C.byref(bp), C.c_void_p(d_energy.data_ptr())))

        ls_ms = timed(rebuild, max(3, args.steps // 2), 2)
        line_search = {"ms_per_rebuild": ls_ms, "what": "collisions_build_from_candidates_dev + barrier_energy_dev at x0 + alpha/2 dx",
                       "resident_candidates": ls_cand, "collisions": list(ls_counts), "alpha": alpha}
        del dX

    # ---- end to end through the host-buffer C ABI (pinned inputs, host results)
    hV0 = torch.from_numpy(np.asfortranarray(V0).T.copy()).pin_memory()
    hV1 = torch.from_numpy(np.asfortranarray(V1).T.copy()).pin_memory()
    h_grad = torch.zeros(3 * nV, dtype=torch.float64).pin_memory()
    hv0p, hv1p = C.c_void_p(hV0.data_ptr()), C.c_void_p(hV1.data_ptr())
    e2e_bytes = {"h2d": 0, "d2h": 0}
    hbuf = {}

    def host_buffers(n):
        if hbuf.get("n", -1) < n:
            hbuf["outer"] = torch.zeros(3 * nV + 1, dtype=torch.int32).pin_memory()
            hbuf["inner"] = torch.zeros(int(n * 1.2) + 1, dtype=torch.int32).pin_memory()
            hbuf["vals"] = torch.zeros(int(n * 1.2) + 1, dtype=torch.float64).pin_memory()
            hbuf["n"] = int(n * 1.2)

    def host_step():
        """N = 1: the host-buffer C ABI — every call takes pinned HOST positions and returns HOST results (energy,
        gradient, the CSR arrays, the step size); the step-size call runs on the second context beside the others"""
        e, st = C.c_double(), C.c_double()
        if ccd_only:
            lib.check(lib.ccd_stepsize(ctx, hv0p, hv1p, nV, 0.0, C.byref(ccd), C.byref(st)))
            e2e_bytes["h2d"], e2e_bytes["d2h"] = 24 * nV * 2, 8
            info["step"], info["energy"] = st.value, 0.0
            return
        job = None
        if stepper.lane is not None:
            job = stepper.lane.submit(lambda: lib.check(lib.ccd_stepsize(stepper.ctx_b, hv0p, hv1p, nV, 0.0, C.byref(ccd), C.byref(st))))
        lib.check(lib.collisions_build(ctx, hv0p, nV, dhat, 0.0, 0, counts))
        lib.check(lib.barrier_energy(ctx, hv0p, nV, C.byref(bp), C.byref(e)))
        lib.check(lib.barrier_gradient(ctx, hv0p, nV, C.byref(bp), C.c_void_p(h_grad.data_ptr())))
        lib.check(lib.barrier_hessian(ctx, hv0p, nV, C.byref(bp), 1, C.byref(nnz)))
        n = nnz.value
        host_buffers(n)
        lib.check(lib.barrier_hessian_fetch(ctx, C.c_void_p(hbuf["outer"].data_ptr()), C.c_void_p(hbuf["inner"].data_ptr()),
                                            C.c_void_p(hbuf["vals"].data_ptr())))
        if job is None:
            lib.check(lib.ccd_stepsize(ctx, hv0p, hv1p, nV, 0.0, C.byref(ccd), C.byref(st)))
        else:
            job.wait()
        e2e_bytes["h2d"] = 24 * nV * 4 + 24 * nV * 2  # V uploaded by build / energy / gradient / hessian + (V0, V1) by ccd
        e2e_bytes["d2h"] = 8 + 24 * nV + 4 * (3 * nV + 1) + 12 * n + 8
        info["step"], info["energy"] = st.value, e.value

    h_small = torch.zeros(2, dtype=torch.float64).pin_memory()

    def sharded_host_step():
        """N > 1: pinned host positions in, host results out around the sharded device step (the all-reduces need
        device buffers): H2D of V0 / V1, the device step with its NCCL all-reduces, D2H of energy, gradient, step size
        and of this rank's Hessian CSR"""
        with torch.cuda.stream(stream):
            dV0.copy_(hV0, non_blocking=True)
            dV1.copy_(hV1, non_blocking=True)
        device_step()
        with torch.cuda.stream(stream):
            h_grad.copy_(d_grad, non_blocking=True)
            h_small[0:1].copy_(d_energy, non_blocking=True)
            h_small[1:2].copy_(d_step, non_blocking=True)
        n = nnz.value
        host_buffers(n)
        lib.check(lib.barrier_hessian_fetch(ctx, C.c_void_p(hbuf["outer"].data_ptr()), C.c_void_p(hbuf["inner"].data_ptr()),
                                            C.c_void_p(hbuf["vals"].data_ptr())))
        e2e_bytes["h2d"] = 24 * nV * 2
        e2e_bytes["d2h"] = 8 + 24 * nV + 4 * (3 * nV + 1) + 12 * n + 8
        info["step"], info["energy"] = float(h_small[1]), float(h_small[0])

    ms_e2e = timed(host_step if world == 1 else sharded_host_step, max(2, args.steps // 2), 1)
    sampler.stop_flag = True

    # ---- roofline denominators measured on this device: HBM copy bandwidth (MEASURED_PEAKS.json, driver-written; the
    # library's own copy kernel beside it) and the FP64 FMA rate (not in MEASURED_PEAKS.json: measured here)
    peak, peak_kind = load_peaks()
    fp64_peak, copy_gbs = C.c_double(), C.c_double()
    lib.check(lib.measure_fp64_peak(ctx, 5, C.byref(fp64_peak)))
    lib.check(lib.measure_copy_bandwidth(ctx, 1 << 30, 5, C.byref(copy_gbs)))

    # ---- rooflines (algorithmic bytes / flops per launch: DESIGN.md §4); kernel times = CUDA events around
    # the kernel on its own stream in the extra recorded step ("k:" timers of the library)
    stages = {k: float(np.sum(v)) for k, v in stage_acc.items()}  # stages that run twice (static + swept) add up
    kernels_ms = {k[2:]: v for k, v in stages.items() if k.startswith("k:")}
    stages = {k: v for k, v in stages.items() if not k.startswith("k:")}
    ncoll = info.get("collisions", [0, 0, 0, 0])
    npts = (2, 3, 4, 4)
    nitems = sum(c * n * n for c, n in zip(ncoll, npts))
    ninc = sum(c * n for c, n in zip(ncoll, npts))
    nnz_ = info.get("nnz", 0) or 0
    traffic = ncu_traffic(args.workload)
    share = 1.0
    if world > 1 and not args.additive_hessian:
        # row-block mode: every rank holds the FULL set and assembles its row block, i.e. about 1 / world of the items and
        # of the block rows; rank 0's share is what its kernels moved (ncu traffic is a single-GPU capture)
        share, traffic = 1.0 / world, None
    cs = info.get("static_candidates") or [0, 0, 0, 0]
    cc = info.get("ccd_candidates") or [0, 0, 0, 0]

    def roof(kernel, ms, alg_bytes, ncu_names, note, flops=None, bound="hbm", fp64_inst=None):
        if not ms or not alg_bytes:
            return None
        ach = alg_bytes / (ms * 1e-3) / 1e9
        hit = [v for k, v in (traffic or {}).items() if any(k.startswith(n) for n in ncu_names)]  # names carry template arguments
        t = sum(hit) if hit and len(hit) >= len(ncu_names) else None
        r = {"kernel": kernel, "bound": bound, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": t,
             "peak_source": (peak_kind + " (MEASURED_PEAKS.json copy bandwidth)") if peak_kind == "measured" else "fallback",
             "algorithmic_bytes": int(alg_bytes), "kernel_ms": ms, "note": note}
        if flops:
            tf = flops / (ms * 1e-3) / 1e12
            r["fp64"] = {"achieved": tf, "peak": fp64_peak.value, "unit": "TFLOP/s", "frac": tf / fp64_peak.value if fp64_peak.value else None,
                         "algorithmic_flops": int(flops), "peak_source": "measured here (ipcb_measure_fp64_peak: 8 FMA chains per thread)"}
            if fp64_inst:
                # what binds an FP64 kernel is the ISSUE rate of the FP64 pipe: a DMUL or a DADD takes the slot of a DFMA.
                # achieved = FP64 thread instructions / s against the measured FMA issue rate (peak TFLOP/s / 2)
                ti, pk = fp64_inst / (ms * 1e-3) / 1e12, fp64_peak.value / 2
                r["fp64"].update({"issue_achieved": ti, "issue_peak": pk, "issue_unit": "T FP64 thread-inst/s", "issue_frac": ti / pk if pk else None})
            if bound == "fp64":
                r["achieved"], r["peak"], r["unit"], r["frac"] = tf, fp64_peak.value, "TFLOP/s", r["fp64"]["frac"]
                if fp64_inst:
                    r["achieved"], r["peak"], r["unit"], r["frac"] = (r["fp64"]["issue_achieved"], r["fp64"]["issue_peak"],
                                                                      "T FP64 thread-inst/s", r["fp64"]["issue_frac"])
        return r

    tri = (3, 6, 10, 10)
    # local Hessians: per collision 24 B record + 32 B per stencil point in; 16 B ids + 32 B masks + 8 B per incidence +
    # 72 B per stored (upper-triangular) 3x3 block out.  FP64: FLOPS_HFAST per collision (from the ncu
    # instruction counts of the capture in profiles/)
    # FP64 work of k_hessian_fast per collision, from the ncu instruction counts of profiles/r2_ncu_full_c3_raw.csv
    # (smsp__sass_thread_inst_executed_op_{dfma,dmul,dadd}_pred_on): thread instructions, and flop with an FMA as two
    FP64_INST = (469.0, 2028.0, 3876.0, 3993.0)
    FLOPS_HFAST = (640.0, 2820.0, 5460.0, 5630.0)
    k_ms = lambda name: kernels_ms.get(name)
    hf_names = ("k_hessian_fast<VV>", "k_hessian_fast<EV>", "k_hessian_fast<EE>", "k_hessian_fast<FV>")
    hl_bytes = [c * share * (24 + 32 * n + 16 + 32 + 8 * n + 72 * t) for c, n, t in zip(ncoll, npts, tri)]
    hl_flops = [c * share * f for c, f in zip(ncoll, FLOPS_HFAST)]
    hl_inst = [c * share * f for c, f in zip(ncoll, FP64_INST)]
    hf_ms = sum(k_ms(n) or 0.0 for n in hf_names)
    # numeric pass: every stored (upper-triangular) 72 B block once — the second read of an off-diagonal block, by the column of
    # its other vertex, is an L2 hit in the Morton visiting order and is not counted —, 4 B reference per item, 8 B per unique
    # block (~nnz / 9) in; 12 B per entry out
    hn_bytes = sum(c * share * 72 * t for c, t in zip(ncoll, tri)) + nitems * share * 4 + (nnz_ // 9) * 8 + nnz_ * 12
    # symbolic pass: 8 B incidence + 16 B ids + 8 B masks per incidence in, 4 B per item + 8 B per unique block out
    hs_bytes = ninc * share * 32 + nitems * share * 4 + (nnz_ // 9) * 8
    rooflines = [r for r in (
        roof("k_hessian_fast<VV|EV|EE|FV> (the four kinds, timed one by one)", hf_ms, sum(hl_bytes),
             ["k_hessian_fast<0", "k_hessian_fast<1", "k_hessian_fast<2", "k_hessian_fast<3"],
             "register Jacobi PSD projection in the analytic (3+p)-subspace; upper-triangular block records leave through per-thread TMA bulk stores (cp.async.bulk)",
             flops=sum(hl_flops), bound="fp64", fp64_inst=sum(hl_inst)),
        roof("k_hessian_fast<EE>", k_ms(hf_names[2]), hl_bytes[2], ["k_hessian_fast<2"], "the dominant kernel of the step", flops=hl_flops[2], bound="fp64",
             fp64_inst=hl_inst[2]),
        roof("k_hess_numeric", stages.get("hess_numeric"), hn_bytes, ["k_hess_numeric"], "HBM gather of 72-byte blocks, software-pipelined"),
        roof("k_hess_symbolic", stages.get("hess_symbolic"), hs_bytes, ["k_hess_symbolic"],
             "shared-memory hash + sort per column; instruction / latency bound"),
        roof("radix sort of the (vertex, collision) incidences (cub)", k_ms("radix_sort(incidences)"), ninc * share * 8 * 2 * 4, ["(no per-sort capture)"],
             "3 onesweep passes + histogram over 8-byte keys (IPCB_HESS_RADIX_INCIDENCES; the default is the counting placement below)"),
        # counting placement: the scatter reads and writes every 8-byte incidence once, the per-column sort once more
        roof("k_scatter_incidences + k_sort_columns", (k_ms("place(incidences)") or 0.0) + (k_ms("sort_columns(incidences)") or 0.0), ninc * share * 32.0,
             ["k_scatter_incidences", "k_sort_columns"], "counting placement of the incidences: atomics + register bitonic sorts per column"),
        # classification: 8 B ids + 2 x 8 B edge / 16 B face ids + 4 x 32 B vertices per candidate; ~150 (EE) / 250 (FV) flop
        roof("k_classify<EE>", k_ms("k_classify<EE>"), cs[2] * 152.0, ["k_classify<2"], "one thread per candidate: dependent gathers out of L2",
             flops=cs[2] * 150.0),
        roof("k_classify<FV>", k_ms("k_classify<FV>"), cs[3] * 152.0, ["k_classify<3"], "one thread per candidate: dependent gathers out of L2",
             flops=cs[3] * 250.0),
        # traversal: 40 B per query leaf + 8 B per emitted pair (static + swept passes); the node fetches (64 B each) hit L2
        roof("k_traverse<EE> (static + swept)", k_ms("k_traverse<EE>"), 2 * E.shape[0] * 40.0 + (cs[2] + cc[2]) * 8.0, ["k_traverse<2,2,2"],
             "one thread per query leaf, stack walk: bound by dependent node fetches (L2 latency), not by DRAM"),
        roof("k_traverse<FV> (static + swept)", k_ms("k_traverse<FV>"), 2 * nV * 40.0 + (cs[3] + cc[3]) * 8.0, ["k_traverse4<1,1,3"],
             "one thread per query leaf, stack walk over the 4-wide face tree: bound by dependent node fetches (L2 latency), not by DRAM"),
        # CCD pre-filter (FP32, conservative): 8 B ids + 16 B edge / face ids + 4 x 32 B re-centred float vertices (t0 | t1) per
        # candidate; the vertex table (16 MB) is L2-resident, so these are L2 -> SM bytes, not DRAM bytes; ~250 FP32 flop
        roof("k_ti_filter32", k_ms("k_ti_filter"), (cc[2] + cc[3]) * 152.0, ["k_ti_filter32"],
             "separating-direction test per swept candidate in FP32 with a conservative margin; gathers out of L2 (bytes = L2->SM traffic)"),
    ) if r]
    roof_main = rooflines[0] if rooflines else None
    if roof_main is not None:  # the contract's `roofline` object: the dominant kernel against the MEASURED HBM bandwidth; its FP64 view rides along
        hbm = roof("k_hessian_fast<VV|EV|EE|FV>", hf_ms, sum(hl_bytes), ["k_hessian_fast<0", "k_hessian_fast<1", "k_hessian_fast<2", "k_hessian_fast<3"],
                   roof_main["note"], flops=sum(hl_flops), bound="hbm", fp64_inst=sum(hl_inst))
        if hbm is not None:
            hbm["binding"] = "fp64 (see the fp64 object: the kernel is bound by the FP64 pipe, its HBM fraction is reported as the contract asks)"
            roof_main = hbm

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # the CPU restatement on the SAME workload, every host core, ONE step (C3: ~20-30 s): a reported baseline
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import pyoracle

        oapi = pyoracle.load(fast=True)
        try:
            cores = len(os.sched_getaffinity(0))
        except AttributeError:
            cores = os.cpu_count() or 1
        oapi.set_num_threads(cores)
        omesh = oapi.CollisionMesh(V0, E, F)
        t = time.perf_counter()
        cpu_out = cpu_step(oapi, omesh, V0, V1, dhat, ccd_only)
        cpu_ms = (time.perf_counter() - t) * 1e3
        cpu = {"value": cpu_ms, "unit": "ms", "cores": cores, "kind": "port",
               "sample": "the full workload (%d triangles), one step, no warm-up" % F.shape[0],
               "collisions": cpu_out[4], "step": cpu_out[3]}
        del omesh, cpu_out

    if rank == 0:
        out = {
            "metric": METRIC, "value": ms_dev, "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": make_config(desc, full_spec, world, args.additive_hessian),
            "e2e": None if ms_e2e is None else {"value": ms_e2e, "unit": "ms", "h2d_bytes_per_step": e2e_bytes["h2d"],
                                                "d2h_bytes_per_step": e2e_bytes["d2h"]},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "roofline": roof_main,
            "rooflines": rooflines,
            "peaks": {"hbm_gbs": peak, "hbm_source": peak_kind, "copy_kernel_gbs": copy_gbs.value, "fp64_tflops": fp64_peak.value,
                      "fp64_source": "ipcb_measure_fp64_peak on this device"},
            "cpu_baseline": cpu,
            "stages_ms": stages,
            "kernels_ms": kernels_ms,
            "lanes": "two contexts: build + potential || swept broad phase + CCD" if ccd_mesh is not None else "one context, sequential calls",
            "broad_phase": args.broad,
            "line_search_rebuild": line_search,
            "counts": {"collisions_rank0": info.get("collisions"), "shard_collisions_rank0": info.get("shard_collisions"),
                       "hessian_rows_rank0": info.get("rows"), "ccd_candidates_rank0": info.get("ccd_candidates"),
                       "static_candidates": info.get("static_candidates"),
                       "hessian_nnz_rank0": info.get("nnz"), "step": info.get("step"),
                       "energy": info.get("energy")},
        }
        line = json.dumps(out) + "\n"
        if json_fd is None:
            sys.stdout.write(line)
        else:
            os.write(json_fd, line.encode())
    # ---- orderly teardown (no os._exit: the driver's exit hook records the loaded libraries).  Everything that was used
    # on the library's streams (wrapped as torch ExternalStreams) is released BEFORE the contexts — and with them the
    # streams — are destroyed: tensors first, then torch's cached blocks, then NCCL, then the contexts.
    torch.cuda.synchronize()
    stepper.release()
    del stepper, dV0, dV1, d_energy, d_grad, d_step, flush, hV0, hV1, h_grad, h_small, stream
    hbuf.clear()
    import gc

    gc.collect()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    if dist is not None:
        dist.destroy_process_group()
    mesh.close()
    if ccd_mesh is not None:
        ccd_mesh.close()
    sys.stdout.flush()
    sys.stderr.flush()


if __name__ == "__main__":
    main()
