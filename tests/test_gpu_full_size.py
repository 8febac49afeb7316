"""Full-size checks (BASELINE.json configs C3 and C2) of the CUDA path through the C ABI, by properties that do not
need the oracle at that size (it would take minutes per step on CPU):

* the set is reproducible (two builds give identical records) and `build(mesh, V)` == `build(candidates, mesh, V)`;
* every collision is active: 0 < min d^2 < dhat^2;
* momentum: the barrier forces sum to zero; translation invariance: H t = 0 for rigid translations t;
* H is symmetric; the CLAMP-projected H is positive semi-definite along random directions;
* directional finite differences: dE/deps ~ g.p and dg/deps ~ H p (unprojected H);
* row blocks tile the matrix, collision ranges add up (the multi-GPU decomposition on one device);
* a step by the returned step size is collision free; Additive CCD and Tight Inclusion agree on the order of magnitude.

The same properties run on the oracle for a small scene in the CPU suite (`test_properties_hold_for_the_oracle`), so
a failure here points at the CUDA path and not at the property.
"""
import numpy as np
import pytest


def relerr(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    n = np.linalg.norm(b)
    return np.linalg.norm(a - b) / n if n > 0 else np.linalg.norm(a)


def check_properties(api, V0, V1, E, F, dhat, seed=0, sets_twice=True):
    rng = np.random.default_rng(seed)
    nV = V0.shape[0]
    mesh = api.CollisionMesh(V0, E, F)
    B = api.BarrierPotential(dhat, 1.0)
    kinds = ("vv", "ev", "ee", "fv")

    c = api.NormalCollisions()
    c.build(mesh, V0, dhat)
    counts = c.counts()
    assert sum(counts) > 0
    recs = [getattr(c, k + "_collisions") for k in kinds]
    for r in recs:  # canonical order: sorted ids, no duplicates among VV / EV / EE
        key = r.ids[:, 0].astype(np.int64) << 32 | r.ids[:, 1].astype(np.int64)
        assert np.all(np.diff(key) >= 0)
    for r in recs[:3]:
        key = r.ids[:, 0].astype(np.int64) << 32 | r.ids[:, 1].astype(np.int64)
        assert np.all(np.diff(key) > 0)
    assert all(np.all(r.weight > 0) for r in recs)
    dmin2 = c.compute_minimum_distance(mesh, V0)
    assert 0 < dmin2 < dhat * dhat

    if sets_twice:
        def same(other):  # a handle is valid until the next build on its mesh: compare right away
            assert other.counts() == counts
            for k, r in zip(kinds, recs):
                o = getattr(other, k + "_collisions")
                assert np.array_equal(o.ids, r.ids) and np.array_equal(o.weight, r.weight)
                assert np.array_equal(o.dtype, r.dtype) and np.array_equal(o.eps_x, r.eps_x)

        c = api.NormalCollisions()
        c.build(mesh, V0, dhat)
        same(c)
        cand = api.Candidates()
        cand.build(mesh, V0, 0.5 * dhat)
        c = api.NormalCollisions()
        c.build(cand, mesh, V0, dhat)
        same(c)

    # evaluate slightly off the build point (keeps every distance positive: the gap is ~0.5 dhat)
    X = V0 + 0.01 * dhat * np.sin(np.arange(V0.size).reshape(V0.shape))
    e = B(c, mesh, X)
    g = B.gradient(c, mesh, X)
    assert e > 0 and np.all(np.isfinite(g))
    G = g.reshape(nV, 3)
    assert np.all(np.abs(G.sum(axis=0)) <= 1e-11 * np.abs(G).sum(axis=0).max())  # momentum

    H0 = B.hessian(c, mesh, X)
    H1 = B.hessian(c, mesh, X, api.PSDProjectionMethod.CLAMP)
    for H in (H0, H1):
        assert H.shape == (3 * nV, 3 * nV)
        scale = np.abs(H.data).max()
        if sets_twice:
            assert abs(H - H.T).max() <= 1e-11 * scale
        else:  # the largest scene: symmetry through two products instead of forming H - H^T (saves minutes of host time)
            v, w = rng.standard_normal(3 * nV), rng.standard_normal(3 * nV)
            assert abs(v @ (H @ w) - w @ (H @ v)) <= 1e-9 * scale * np.sqrt(v @ v) * np.sqrt(w @ w)
        for axis in range(3):  # translation invariance
            t = np.zeros(3 * nV)
            t[axis::3] = 1.0
            assert np.abs(H @ t).max() <= 1e-9 * scale
    for _ in range(3):  # CLAMP: positive semi-definite
        v = rng.standard_normal(3 * nV)
        assert v @ (H1 @ v) >= -1e-9 * np.abs(H1.data).max() * (v @ v)
    assert H1.nnz >= H0.nnz * 0.5  # same block structure (projection can only change which entries are exactly zero)

    # directional finite differences (central), direction scaled to the contact gap
    p = rng.standard_normal(V0.shape)
    eps = 1e-5 * dhat
    ep, em = B(c, mesh, X + eps * p), B(c, mesh, X - eps * p)
    gp, gm = B.gradient(c, mesh, X + eps * p), B.gradient(c, mesh, X - eps * p)
    pf = p.reshape(-1)
    assert abs((ep - em) / (2 * eps) - g @ pf) <= 1e-5 * abs(g @ pf)
    assert relerr((gp - gm) / (2 * eps), H0 @ pf) <= 1e-5

    # the multi-GPU decomposition on one device: collision ranges add up, row blocks tile
    world = 4
    bounds = mesh.balanced_row_blocks(world)
    assert bounds[0] == 0 and bounds[-1] == nV and np.all(np.diff(bounds) > 0)
    e_sum, g_sum, nnz_sum = 0.0, np.zeros_like(g), 0
    Hc = H1.tocsc()
    loads = []
    for r in range(world):
        mesh.set_collision_range(r, world)
        mesh.set_row_block(int(bounds[r]), int(bounds[r + 1]))
        e_sum += B(c, mesh, X)
        g_sum += B.gradient(c, mesh, X)
        T = B.hessian(c, mesh, X, api.PSDProjectionMethod.CLAMP).tocsc()
        lo, hi = 3 * int(bounds[r]), 3 * int(bounds[r + 1])
        assert np.array_equal(T.indptr[lo:hi + 1] - T.indptr[lo], Hc.indptr[lo:hi + 1] - Hc.indptr[lo])
        a, b = T.indptr[lo], T.indptr[hi]
        assert T.indptr[lo] == 0 and T.indptr[-1] == b  # nothing outside the block
        assert np.array_equal(T.indices[a:b], Hc.indices[Hc.indptr[lo]:Hc.indptr[hi]])
        assert relerr(T.data[a:b], Hc.data[Hc.indptr[lo]:Hc.indptr[hi]]) <= 1e-13
        nnz_sum += T.nnz
        loads.append(T.nnz)
    mesh.set_collision_range(0, 1)
    mesh.set_row_block()
    assert nnz_sum == H1.nnz
    assert abs(e_sum - e) <= 1e-12 * e and relerr(g_sum, g) <= 1e-12
    assert max(loads) <= 1.5 * (sum(loads) / world)  # the blocks are balanced

    # CCD
    ti = api.compute_collision_free_stepsize(mesh, V0, V1)
    ac = api.compute_collision_free_stepsize(mesh, V0, V1, narrow_phase_ccd=api.AdditiveCCD())
    assert 0 < ti <= 1 and 0 < ac <= 1
    if ti < 1:
        assert ac < 1 and 0.25 <= ac / ti <= 4.0
        assert api.is_step_collision_free(mesh, V0, V0 + 0.999 * ti * (V1 - V0))
        assert not api.is_step_collision_free(mesh, V0, V1)
    return dict(counts=counts, nnz=H1.nnz, step=ti)


def test_properties_hold_for_the_oracle(oracle, scenes):
    V0, V1, E, F, P = scenes.cloth_stack(3, 24)
    check_properties(oracle, V0, V1, E, F, P["dhat"])


@pytest.mark.gpu
def test_properties_small_scene(cuda, scenes):
    V0, V1, E, F, P = scenes.cloth_stack(3, 24)
    check_properties(cuda, V0, V1, E, F, P["dhat"])


@pytest.mark.gpu
def test_c2_drape_full_size(cuda, scenes):
    """BASELINE config 2: 256 x 256 cloth draped on a 50K-triangle sphere (~180K triangles)"""
    V0, V1, E, F, P = scenes.cloth_on_sphere(256, 160, drape=True)
    assert F.shape[0] > 170_000
    check_properties(cuda, V0, V1, E, F, P["dhat"])


@pytest.mark.gpu
def test_c3_stack_full_size(cuda, scenes):
    """BASELINE config 3 (the metric's configuration): 8 stacked 250 x 250 sheets, 1M triangles, dense edge-edge contact"""
    V0, V1, E, F, P = scenes.cloth_stack(8, 250, gap=0.5)
    assert F.shape[0] == 1_000_000
    out = check_properties(cuda, V0, V1, E, F, P["dhat"], sets_twice=False)
    assert out["counts"][2] > 3_000_000 and out["nnz"] > 100_000_000
