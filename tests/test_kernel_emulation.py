"""Host emulation of the collision-set kernels (CPU suite): the merge kernels and the whole IMPROVED_MAX_APPROX section
of ipc-toolkit_b200/csrc/collisions.cu are cut out of the .cu between their [emu-begin]/[emu-end] tags, compiled with
g++ behind a shim that maps the CUDA built-ins onto one-lane warps (tests/cpp/emu_collisions.cpp), fed with the oracle's
candidates and IPC records, and their output is compared with the oracle's IMPROVED_MAX_APPROX set.  This checks the
kernels' logic without a GPU; `tests/test_zz_improved_max_approx_gpu.py` is the hardware run."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "ipc-toolkit_b200", "csrc")
CUDA_INC = "/usr/local/cuda/include"


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    if not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    out = tmp_path_factory.mktemp("emu")
    src = open(os.path.join(CSRC, "collisions.cu")).read()
    parts = re.findall(r"// \[emu-begin (\w+)\].*?\n(.*?)// \[emu-end \1\]", src, flags=re.S)
    assert [p[0] for p in parts] == ["merge", "improved"]
    open(os.path.join(out, "emu_kernels.inc"), "w").write("\n".join(p[1] for p in parts))
    lib = os.path.join(out, "libemu.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-std=c++17", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-w", "-I", CUDA_INC, "-I", CSRC, "-I", str(out),
                    os.path.join(ROOT, "tests", "cpp", "emu_collisions.cpp"), "-o", lib], check=True)
    return C.CDLL(lib)


def _csr(rows):
    off = np.zeros(len(rows) + 1, np.int32)
    for i, r in enumerate(rows):
        off[i + 1] = off[i] + len(r)
    val = np.array([x for r in rows for x in r] + [0], np.int32)
    return off, val


def _adjacency(nV, E, F, F2E):  # collision_mesh.cpp:247-307
    vv, ve, ev = [set() for _ in range(nV)], [set() for _ in range(nV)], [set() for _ in range(len(E))]
    for i, (a, b) in enumerate(E):
        vv[a].add(b), vv[b].add(a), ve[a].add(i), ve[b].add(i)
    for i in range(len(F)):
        for j in range(3):
            ev[F2E[i, j]].add(int(F[i, (j + 2) % 3]))
    boundary = np.ones(max(nV, 1), np.uint8)
    for i, (a, b) in enumerate(E):
        if len(ev[i]) >= 2:
            boundary[a] = boundary[b] = 0
    srt = lambda rows: [sorted(r) for r in rows]
    return _csr(srt(vv)), _csr(srt(ve)), _csr(srt(ev)), boundary, max([len(r) for r in ve] + [1])


def _pad4(a, dtype):
    out = np.zeros((len(a), 4), dtype)
    out[:, :a.shape[1]] = a
    return np.ascontiguousarray(out)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("name", ["stack", "drape", "dense", "soup", "codim"])
@pytest.mark.parametrize("area", [False, True])
def test_improved_max_approx_kernels_on_the_host(emu, oracle, scenes, name, area):
    if name == "codim":  # edge-vertex candidates only exist between codimensional edges and vertices
        V0 = np.array([[0, 0, 0], [1, 0, 0], [0, 0, -1], [-1, 0, 0], [0, 0, 1], [0, 1, 0], [0, 2, 0], [0, 3, 0]], float)
        E, F, dhat, dmin = np.array([[0, 1], [0, 2], [0, 3], [0, 4]], np.int32), np.zeros((0, 3), np.int32), 0.25, 0.8
    else:
        V0, V1, E, F, P = {"stack": lambda: scenes.cloth_stack(3, 12), "drape": lambda: scenes.cloth_on_sphere(24, 12, drape=True),
                           "dense": lambda: scenes.dense_sheet(8, 2.0), "soup": lambda: scenes.random_soup(60, seed=7)}[name]()
        dhat, dmin = P["dhat"], 0.0
    T = oracle.NormalCollisions.CollisionSetType
    mesh = oracle.CollisionMesh(V0, E, F)
    nV, nE, nF = len(V0), len(E), len(F)
    F2E = mesh.faces_to_edges() if nF else np.zeros((0, 3), np.int32)
    cand = oracle.Candidates()
    cand.build(mesh, V0, 0.5 * (dhat + dmin))
    cands = [np.ascontiguousarray(c, np.int32).reshape(-1, 2) for c in (cand.vv_candidates, cand.ev_candidates, cand.ee_candidates, cand.fv_candidates)]

    def records(t):
        c = oracle.NormalCollisions()
        c.set_use_area_weighting(area)
        c.set_collision_set_type(t)
        c.build(cand, mesh, V0, dhat, dmin)
        return [getattr(c, k + "_collisions") for k in ("vv", "ev", "ee", "fv")]

    ipc, want = records(T.IPC), records(T.IMPROVED_MAX_APPROX)
    (vvo, vvv), (veo, vev), (evo, evv), boundary, max_ve = _adjacency(nV, E.tolist(), F, F2E)
    X4, R4 = _pad4(V0, np.float64), _pad4(V0, np.float64)
    E2, F4, F2E4 = np.ascontiguousarray(E, np.int32), _pad4(F, np.int32), _pad4(np.asarray(F2E), np.int32)
    va, ea = mesh.vertex_areas(), mesh.edge_areas()
    ncand = (C.c_int64 * 4)(*[len(c) for c in cands])
    cand_p = (C.c_void_p * 4)(*[_ptr(c) for c in cands])
    rec_ids = [np.ascontiguousarray(r.ids, np.int32) for r in ipc]
    rec_w = [np.ascontiguousarray(r.weight, np.float64) for r in ipc]
    nrec = (C.c_int64 * 4)(*[len(r) for r in rec_ids])
    ids_p, w_p = (C.c_void_p * 4)(*[_ptr(r) for r in rec_ids]), (C.c_void_p * 4)(*[_ptr(r) for r in rec_w])
    ee_eps, ee_dt = np.ascontiguousarray(ipc[2].eps_x), np.ascontiguousarray(ipc[2].dtype, np.uint8)
    cap = sum(len(r) for r in rec_ids) + 6 * sum(len(c) for c in cands) * (max_ve + 1) + 16
    out_ids = [np.zeros((cap, 2), np.int32) for _ in range(4)]
    out_w = [np.zeros(cap) for _ in range(4)]
    out_eps, out_dt = np.zeros(cap), np.zeros(cap, np.uint8)
    out_count = (C.c_int64 * 4)()
    oi, ow = (C.c_void_p * 4)(*[_ptr(a) for a in out_ids]), (C.c_void_p * 4)(*[_ptr(a) for a in out_w])
    emu.emu_improved_build.argtypes = [C.c_int] * 3 + [C.c_void_p] * 7 + [C.c_void_p] * 7 + [C.c_void_p] * 7 + [C.c_int, C.c_double, C.c_int] + [C.c_void_p] * 5
    rc = emu.emu_improved_build(nV, nE, nF, _ptr(X4), _ptr(R4), _ptr(E2), _ptr(F4), _ptr(F2E4), _ptr(va), _ptr(ea),
                                C.cast(ncand, C.c_void_p), C.cast(cand_p, C.c_void_p), C.cast(nrec, C.c_void_p), C.cast(ids_p, C.c_void_p),
                                C.cast(w_p, C.c_void_p), _ptr(ee_eps), _ptr(ee_dt), _ptr(vvo), _ptr(vvv), _ptr(veo), _ptr(vev), _ptr(evo),
                                _ptr(evv), _ptr(boundary), int(max_ve), (dmin + dhat) ** 2, int(area), C.cast(out_count, C.c_void_p),
                                C.cast(oi, C.c_void_p), C.cast(ow, C.c_void_p), _ptr(out_eps), _ptr(out_dt))
    assert rc == 0
    assert sum(len(r.ids) for r in want) > 0
    changed = False
    for kind in range(4):
        n = out_count[kind]
        got_ids, got_w = out_ids[kind][:n], out_w[kind][:n]
        w = want[kind]
        scale = max(np.abs(w.weight).max(), 1e-300) if len(w.weight) else 1.0
        ka, kb = np.abs(got_w) > 1e-12 * scale, np.abs(w.weight) > 1e-12 * scale  # see test_zz_improved_max_approx_gpu.py
        assert np.array_equal(got_ids[ka], w.ids[kb]), "kind %d" % kind
        assert np.allclose(got_w[ka], w.weight[kb], rtol=1e-12, atol=0)
        if kind == 2:
            assert np.array_equal(out_dt[:n][ka], w.dtype[kb]) and np.array_equal(out_eps[:n][ka], w.eps_x[kb])
        changed |= len(w.ids) != len(ipc[kind].ids) or not np.array_equal(w.weight, ipc[kind].weight)
    assert changed or name == "soup"  # the corrections did something
