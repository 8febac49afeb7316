"""Generates tests/golden/*.npz — run in the build container:  python tests/golden/make_golden.py [--full c2 c3 c5]

1. reference_kats.npz: the transcribed reference test vectors of tests/kats.py frozen as arrays (the
   analytic expectations come from the reference's own test sources, cited in kats.py);
2. scene_<name>.npz: outputs of the CPU oracle (oracle/, test infrastructure) on the seeded synthetic
   scenes: candidate / collision sets, weights, energy, gradient, Hessian (CSC) for the three
   projection modes, step sizes.  The upstream library cannot be built or imported here (SURVEY §8c:
   no Eigen / TBB / Tight-Inclusion in the image), so these are oracle outputs, not upstream outputs;
   the oracle itself is pinned by tests/test_reference_kats.py.
3. --full: digest_<config>.npz — tests/digest.py digests of one contact step of the ORACLE on the BASELINE.json
   configurations at FULL size (C2 180K, C3 1M, C5 2M triangles): counts + SHA-256 of every index array, norms /
   samples / random projections of every value array, step sizes and the derived Tight-Inclusion time tolerance
   (tests/ccd_tolerance.py).  tests/test_gpu_full_size.py compares the CUDA path against them on the GPU box.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import kats  # noqa: E402
import pyoracle  # noqa: E402
import ipctk_b200  # noqa: E402

SCENES = {
    "c1_small": lambda s: s.cloth_on_sphere(24, 12),
    "drape": lambda s: s.cloth_on_sphere(20, 12, drape=True),
    "stack": lambda s: s.cloth_stack(3, 14),
    "soup": lambda s: s.random_soup(60, seed=7),
}


FULL = {
    # name: (scene, Hessian projection modes in the digest, row blocks the ORACLE assembles the Hessian in: bounded memory)
    "c2": (lambda s: s.cloth_on_sphere(256, 160, drape=True), (0, 1, 2), 1),
    "c3": (lambda s: s.cloth_stack(8, 250, gap=0.5), (1,), 2),
    "c5": (lambda s: s.cloth_stack(16, 250, gap=0.2), (1,), 12),
}


def full_digests(names):
    import time

    import digest
    from ccd_tolerance import StepTolerance

    o = pyoracle.load()
    scenes = ipctk_b200._pkg.scenes
    for name in names:
        make, modes, blocks = FULL[name]
        V0, V1, E, F, P = make(scenes)
        t = time.time()
        d = digest.step_digest(o, V0, V1, E, F, P["dhat"], modes=modes, hess_blocks=blocks)
        T = StepTolerance(o, V0, V1, E, F)
        tol, v, w = T.tolerance(d["step_ti"])
        d["step_ti_tolerance"], d["step_ti_closing_speed"], d["step_ti_box_width"] = tol, v, w
        digest.save(os.path.join(HERE, "digest_%s.npz" % name), d)
        print(name, "triangles", F.shape[0], "collisions", d["counts"].tolist(), "nnz", d["h1_nnz"], "energy", d["energy"], "step", d["step_ti"],
              d["step_accd"], "tol", tol, "oracle seconds %.0f" % (time.time() - t), flush=True)


def main():
    if "--full" in sys.argv:
        return full_digests(sys.argv[sys.argv.index("--full") + 1:])
    o = pyoracle.load()
    scenes = ipctk_b200._pkg.scenes
    # ---- 1. reference vectors
    pt = kats.point_triangle_type_cases()
    ee = kats.edge_edge_not_ea_eb_cases() + kats.edge_edge_ea_eb_cases()
    np.savez_compressed(
        os.path.join(HERE, "reference_kats.npz"),
        pt_x=np.array([np.concatenate(c[:4]) for c in pt]), pt_type=np.array([c[4] for c in pt], np.int8),
        ee_x=np.array([np.concatenate(c[:4]) for c in ee]), ee_d2=np.array([c[5] for c in ee]),
        ee_ok=np.array([[t in c[4] for t in range(9)] for c in ee]),
        ptd_x=np.array([np.concatenate(c[:4]) for c in kats.point_triangle_distance_cases()]),
        ptd_cp=np.array([c[4] for c in kats.point_triangle_distance_cases()]),
        eed_x=np.array([np.concatenate(c[:4]) for c in kats.edge_edge_distance_grid() + kats.edge_edge_degenerate_cases()]),
        eed_d2=np.array([c[4] for c in kats.edge_edge_distance_grid() + kats.edge_edge_degenerate_cases()]))
    # ---- 2. oracle outputs on seeded scenes
    for name, make in SCENES.items():
        V0, V1, E, F, P = make(scenes)
        dhat = P["dhat"]
        mesh = o.CollisionMesh(V0, E, F)
        cand = o.Candidates()
        cand.build(mesh, V0, 0.5 * dhat)
        out = dict(V0=V0, V1=V1, E=E, F=F, dhat=dhat, ee_cand=np.asarray(cand.ee_candidates), fv_cand=np.asarray(cand.fv_candidates))
        c = o.NormalCollisions()
        c.build(mesh, V0, dhat)
        for k in ("vv", "ev", "ee", "fv"):
            s = getattr(c, k + "_collisions")
            out[k + "_ids"], out[k + "_w"] = s.ids, s.weight
        out["ee_dtype"], out["ee_eps"] = c.ee_collisions.dtype, c.ee_collisions.eps_x
        B = o.BarrierPotential(dhat, 1.0)
        X = V0 + 0.02 * dhat * np.sin(np.arange(V0.size).reshape(V0.shape))
        out["X"] = X
        out["energy"] = B(c, mesh, X)
        out["gradient"] = B.gradient(c, mesh, X)
        for mode in (0, 1, 2):
            H = B.hessian(c, mesh, X, mode)
            out["h%d_indptr" % mode], out["h%d_indices" % mode], out["h%d_data" % mode] = H.indptr, H.indices, H.data
        out["step_ti"] = o.compute_collision_free_stepsize(mesh, V0, V1)
        out["step_accd"] = o.compute_collision_free_stepsize(mesh, V0, V1, narrow_phase_ccd=o.AdditiveCCD())
        np.savez_compressed(os.path.join(HERE, "scene_%s.npz" % name), **out)
        print(name, c.counts(), out["energy"], out["step_ti"], out["step_accd"])


if __name__ == "__main__":
    main()
