"""Digest of one contact step on a scene — the full-size parity evidence (BASELINE configs C2 / C3 / C5).

`step_digest(api, ...)` drives either library (the CUDA product or the CPU oracle) through the same API calls and
reduces every output to something small enough to commit under tests/golden/ and strong enough to pin parity:

* index / byte data (candidate sets, collision ids, edge-edge distance types and eps_x, Hessian outer / inner arrays):
  counts + SHA-256 of the canonical arrays  -> compared for equality (bit-exact bar);
* floating-point data (energy, gradient, Hessian values): norms, a strided subsample and products with seeded random
  vectors -> compared to 1e-10 relative (north_star bar);
* step sizes: Tight Inclusion and Additive CCD.

`compare_digests(got, want)` applies the bars and returns a list of human-readable mismatches.
"""
import hashlib

import numpy as np

RTOL = 1e-10
KINDS = ("vv", "ev", "ee", "fv")


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def eval_point(V0, dhat):
    """positions slightly off the build point (keeps every distance positive: the gaps are >= 0.2 dhat)"""
    return V0 + 0.01 * dhat * np.sin(np.arange(V0.size).reshape(V0.shape))


SAMPLE_STRIDE = 65536


def _hessian_digest(api, d, p, B, c, mesh, X, mode, R, blocks):
    """Hessian part of the digest.  blocks > 1 assembles the matrix in that many balanced row blocks, one after the other
    (ctx_set_row_block: the block matrices tile the global compressed-column arrays, so hashing / summing them block by
    block gives the digest of the whole matrix) — what lets the CPU oracle digest the 2M-triangle scene in bounded memory."""
    nV = mesh.num_vertices()
    bounds = mesh.balanced_row_blocks(blocks) if blocks > 1 else np.array([0, nV])
    h_idx, nnz, sq, ab = hashlib.sha256(), 0, 0.0, 0.0
    indptr = np.zeros(3 * nV + 1, np.int64)
    sample = []
    Y = np.zeros((R.shape[0], 3 * nV))
    for b in range(len(bounds) - 1):
        lo, hi = 3 * int(bounds[b]), 3 * int(bounds[b + 1])
        if blocks > 1:
            mesh.set_row_block(int(bounds[b]), int(bounds[b + 1]))
        H = B.hessian(c, mesh, X, mode)
        a, e = int(H.indptr[lo]), int(H.indptr[hi])
        assert a == 0 and e == H.nnz, "entries outside the row block"
        indptr[lo:hi + 1] = H.indptr[lo:hi + 1].astype(np.int64) + nnz
        h_idx.update(np.ascontiguousarray(H.indices, np.int32).tobytes())
        first = (-nnz) % SAMPLE_STRIDE
        sample.append(H.data[first::SAMPLE_STRIDE].copy())
        sq += float(H.data @ H.data)
        ab += float(np.abs(H.data).sum())
        for i in range(R.shape[0]):
            Y[i] += H @ R[i]
        nnz += int(H.nnz)
        del H
    if blocks > 1:
        mesh.set_row_block()
    indptr[3 * nV:] = nnz
    d[p + "nnz"] = nnz
    d[p + "indptr_sha"] = _sha(indptr.astype(np.int32))
    d[p + "indices_sha"] = h_idx.hexdigest()
    d[p + "data_norm"] = sq ** 0.5
    d[p + "data_abssum"] = ab
    d[p + "data_sample"] = np.concatenate(sample)
    d[p + "Hr_norm"] = np.linalg.norm(Y, axis=1)
    d[p + "rHr"] = np.array([[R[i] @ Y[j] for j in range(R.shape[0])] for i in range(R.shape[0])])
    d[p + "Hr_sample"] = Y[:, :: max(1, 3 * nV // 2048)].copy()


def step_digest(api, V0, V1, E, F, dhat, modes=(1,), ccd=True, candidates=True, seed=12345, hess_blocks=1):
    rng = np.random.default_rng(seed)
    nV = V0.shape[0]
    mesh = api.CollisionMesh(V0, E, F)
    d = {"nV": nV, "nE": int(E.shape[0]), "nF": int(F.shape[0]), "dhat": float(dhat)}
    if candidates:
        cand = api.Candidates()
        cand.build(mesh, V0, 0.5 * dhat)
        for k in ("ee", "fv"):
            a = np.asarray(getattr(cand, k + "_candidates"), np.int32)
            d["cand_%s_n" % k] = int(a.shape[0])
            d["cand_%s_sha" % k] = _sha(a)
        del cand
    c = api.NormalCollisions()
    c.build(mesh, V0, dhat)
    d["counts"] = np.asarray(c.counts(), np.int64)
    for k in KINDS:
        r = getattr(c, k + "_collisions")
        d["coll_%s_ids_sha" % k] = _sha(np.asarray(r.ids, np.int32))
        d["coll_%s_wsum" % k] = float(np.sum(r.weight))
    r = c.ee_collisions
    d["coll_ee_dtype_sha"] = _sha(np.asarray(r.dtype, np.uint8))
    d["coll_ee_eps_sha"] = _sha(np.asarray(r.eps_x, np.float64))
    d["coll_ee_dtype_hist"] = np.bincount(np.asarray(r.dtype, np.uint8), minlength=10).astype(np.int64)
    B = api.BarrierPotential(dhat, 1.0)
    X = eval_point(V0, dhat)
    d["min_distance_sqr"] = float(c.compute_minimum_distance(mesh, X))
    d["energy"] = float(B(c, mesh, X))
    g = B.gradient(c, mesh, X)
    R = rng.standard_normal((3, 3 * nV))
    d["grad_norm"] = float(np.linalg.norm(g))
    d["grad_proj"] = R @ g
    d["grad_sample"] = g[:: max(1, g.size // 4096)].copy()
    for mode in modes:
        _hessian_digest(api, d, "h%d_" % mode, B, c, mesh, X, api.PSDProjectionMethod(mode), R, hess_blocks)
    if ccd:
        d["step_ti"] = float(api.compute_collision_free_stepsize(mesh, V0, V1))
        d["step_accd"] = float(api.compute_collision_free_stepsize(mesh, V0, V1, narrow_phase_ccd=api.AdditiveCCD()))
    return d


def _rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    n = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / n) if n > 0 else float(np.linalg.norm(a))


def compare_digests(got, want, step_tol=None, rtol=RTOL):
    """[] when `got` meets the parity bars against `want`.  step_tol: derived time tolerance of the Tight-Inclusion step
    (tests/ccd_tolerance.py); None skips the step sizes."""
    bad = []
    for k in sorted(want.keys()):
        if k not in got:
            continue
        w, g = want[k], got[k]
        if k.startswith("step_"):
            continue
        if k.endswith("_sha") or isinstance(w, str):
            if str(g) != str(w):
                bad.append("%s differs (bit-exact bar)" % k)
        elif np.asarray(w).dtype.kind in "iu":
            if not np.array_equal(np.asarray(g), np.asarray(w)):
                bad.append("%s: %s != %s" % (k, np.asarray(g).tolist(), np.asarray(w).tolist()))
        else:
            tol = 1e-12 if k.endswith("wsum") or k == "dhat" else rtol
            e = _rel(g, w)
            if not e <= tol:
                bad.append("%s: relative error %.3e > %.1e" % (k, e, tol))
    if step_tol is not None and "step_ti" in want and "step_ti" in got:
        if not abs(got["step_ti"] - want["step_ti"]) <= step_tol:
            bad.append("step_ti: %.9g vs %.9g, derived tolerance %.3e" % (got["step_ti"], want["step_ti"], step_tol))
        if not abs(got["step_accd"] - want["step_accd"]) <= 1e-9 * max(want["step_accd"], 1e-12):
            bad.append("step_accd: %.12g vs %.12g" % (got["step_accd"], want["step_accd"]))
    return bad


def save(path, d):
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in d.items()})


def load(path):
    z = np.load(path, allow_pickle=False)
    out = {}
    for k in z.files:
        v = z[k]
        out[k] = v.item() if v.ndim == 0 else v
    return out
