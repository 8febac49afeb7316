"""Pins the CPU oracle against the reference's own known-answer / property tests, transcribed in
tests/kats.py with the reference file:line of every case.  CPU only; the GPU versions of the same vectors are in
tests/test_gpu_kats.py."""
import ctypes as C

import numpy as np
import pytest

import kats


def _p(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.c_void_p)


def dtype_of(o, kind, pts):
    a, p = _p(np.concatenate([np.asarray(x, float) for x in pts]))
    return o.cdll.ipco_unit_distance_type(kind, p)


def distance(o, kind, pts, dtype=-1, derivs=False):
    a, p = _p(np.concatenate([np.asarray(x, float) for x in pts] + [np.zeros(12 - 3 * len(pts))]))
    val = C.c_double()
    g, gp = _p(np.zeros(12))
    h, hp = _p(np.zeros(144))
    rc = o.cdll.ipco_unit_distance(kind, p, dtype, C.byref(val), gp, hp)
    assert rc == 0
    return (val.value, g, h.reshape(12, 12).T) if derivs else val.value


# ---- distance types -------------------------------------------------------------------------------
def test_point_edge_distance_type(oracle):
    for p, e0, e1, ok in kats.point_edge_type_cases():
        assert dtype_of(oracle, 1, (p, e0, e1)) in ok


def test_point_triangle_distance_type(oracle):
    for p, t0, t1, t2, expected in kats.point_triangle_type_cases():
        assert dtype_of(oracle, 3, (p, t0, t1, t2)) == expected, (p, expected)


def test_edge_edge_distance_type_and_distance(oracle):
    for gen in (kats.edge_edge_not_ea_eb_cases, kats.edge_edge_ea_eb_cases, kats.edge_edge_parallel_cases):
        for a0, a1, b0, b1, ok, d2 in gen():
            if ok is not None:
                assert dtype_of(oracle, 2, (a0, a1, b0, b1)) in ok
            d = distance(oracle, 2, (a0, a1, b0, b1))
            assert d == pytest.approx(d2, abs=1e-12, rel=1e-9)
            if ok is None:  # test_edge_edge.cpp:206-211: AUTO is never larger than any explicit vertex / edge type
                for t in range(8):
                    assert d <= distance(oracle, 2, (a0, a1, b0, b1), t) * (1 + 1e-9) + 1e-15


def test_edge_edge_coplanar_regression(oracle):
    for a0, a1, b0, b1, required, expected in kats.edge_edge_coplanar_regression():
        t = dtype_of(oracle, 2, (a0, a1, b0, b1))
        assert t != kats.EA_EB
        if required is not None:
            assert t == required
        d = distance(oracle, 2, (a0, a1, b0, b1))
        assert d > 0 and np.isfinite(d)
        if expected is not None:
            assert d == pytest.approx(expected, abs=1e-10)


# ---- distances -----------------------------------------------------------------------------------
def test_point_triangle_distance(oracle):
    for p, t0, t1, t2, cp in kats.point_triangle_distance_cases():
        assert distance(oracle, 3, (p, t0, t1, t2)) == pytest.approx(float(np.sum((p - cp) ** 2)), abs=1e-12, rel=1e-12)


def test_edge_edge_distance(oracle):
    for e00, e01, e10, e11, d2 in kats.edge_edge_distance_grid() + kats.edge_edge_degenerate_cases():
        assert distance(oracle, 2, (e00, e01, e10, e11)) == pytest.approx(d2, abs=1e-12, rel=1e-12)


def _fd_grad(f, x, h=1e-7):
    g = np.zeros_like(x)
    for i in range(x.size):
        e = np.zeros_like(x)
        e[i] = h * max(1.0, abs(x[i]))
        g[i] = (f(x + e) - f(x - e)) / (2 * e[i])
    return g


def test_point_point_and_point_edge_distance(oracle):
    """distance/test_point_point.cpp:19-33 (aligned / diagonal vectors of length d -> d^2) and test_point_edge.cpp:21-37,
    61-97 (p at height d over a long edge; p = e0 + alpha (e1 - e0) + d n over random edges: d^2 inside, end-point distances
    outside)"""
    ds = (-10, -1, -1e-12, 0, 1e-12, 1, 10)
    for d in ds:
        assert distance(oracle, 0, ([0, 0, 0], [d, 0, 0])) == pytest.approx(d * d)
        assert distance(oracle, 0, ([0, 0, 0], np.ones(3) / np.sqrt(3) * d)) == pytest.approx(d * d)
        assert distance(oracle, 1, ([0, d, 0], [-10, 0, 0], [10, 0, 0])) == pytest.approx(d * d)
    rng = np.random.default_rng(5)
    for alpha in np.arange(-1.0, 2.0, 0.1):
        for d in np.arange(-10.0, 10.0, 1.0):
            for _ in range(4):
                e0, e1 = rng.uniform(-1, 1, 3), rng.uniform(-1, 1, 3)
                n = np.cross(e1 - e0, [1.0, 0, 0])
                n /= np.linalg.norm(n)
                pnt = (e1 - e0) * alpha + e0 + d * n
                want = np.linalg.norm(e0 - pnt) if alpha < 0 else (np.linalg.norm(e1 - pnt) if alpha > 1 else d)
                got = distance(oracle, 1, (pnt, e0, e1))
                # alpha within rounding of 0 or 1 may be classified either way: both answers agree to rounding there
                assert got == pytest.approx(want * want, rel=1e-9, abs=1e-12), (alpha, d)


def test_point_plane_and_line_line_distance(oracle):
    """distance/test_point_plane.cpp:17-31 (plane y = y_plane through t0, t1, t2: (y - y_plane)^2) and test_line_line.cpp:37-46
    (perpendicular lines at heights ya, yb: (ya - yb)^2), i.e. the point-triangle / edge-edge distances with the interior types"""
    rng = np.random.default_rng(9)
    for _ in range(200):
        x, y, z, yp = rng.uniform(-10, 10, 4)
        d = distance(oracle, 3, ([x, y, z], [-1, yp, 0], [1, yp, -1], [1, yp, 0]), kats.P_T)
        assert d == pytest.approx((y - yp) ** 2, rel=1e-12, abs=1e-12)
        ya, yb = rng.uniform(-100, 100, 2)
        d = distance(oracle, 2, ([-1, ya, 0], [1, ya, 0], [0, yb, -1], [0, yb, 1]), kats.EA_EB)
        assert d == pytest.approx((ya - yb) ** 2, rel=1e-12, abs=1e-12)


def test_distance_derivatives_match_finite_differences(oracle):
    """the reference checks every gradient / Hessian against finite differences (distance/test_*.cpp "gradient"/"hessian"
    cases); same check on random stencils for every kind and every explicit distance type"""
    rng = np.random.default_rng(3)
    for kind, npts, ntypes in ((0, 2, 1), (1, 3, 3), (2, 4, 9), (3, 4, 7)):
        for t in range(ntypes):
            for _ in range(3):
                x = rng.uniform(-1, 1, 3 * npts)
                dt = -1 if kind == 0 else t
                val, g, H = distance(oracle, kind, x.reshape(-1, 3), dt, derivs=True)
                f = lambda y: distance(oracle, kind, y.reshape(-1, 3), dt)
                fg = _fd_grad(f, x)
                assert np.allclose(g[: 3 * npts], fg, rtol=1e-5, atol=1e-6)
                fH = np.array([_fd_grad(lambda y, i=i: distance(oracle, kind, y.reshape(-1, 3), dt, derivs=True)[1][i], x) for i in range(3 * npts)])
                assert np.allclose(H[: 3 * npts, : 3 * npts], fH, rtol=1e-4, atol=1e-5)


# ---- mollifier, barrier, PSD, Morton -----------------------------------------------------------------------
def test_edge_edge_mollifier(oracle):
    """distance/test_edge_edge_mollifier.cpp:18-292"""
    cd = oracle.cdll
    perp = np.array([[-1.0, 0, 0], [1, 0, 0], [0, -1, 0], [0, 1, 0]])
    almost = np.array([[-1.0, 0, 0], [1, 0, 0], [-1, 1e-9, 0], [1, -1e-9, 0]])
    rest = np.array([[0.0, 0, 0], [1, 0, 0], [0, 0, 0], [0, 1, 0]])

    def moll(x, eps):
        a, p = _p(x)
        out, op = _p(np.zeros(2))
        g, gp = _p(np.zeros(12))
        h, hp = _p(np.zeros(144))
        cd.ipco_unit_mollifier(p, eps, op, gp, hp)
        return out[0], out[1], g, h.reshape(12, 12).T

    assert moll(perp, 1.0)[0] == pytest.approx(16)  # cross squared norm, :22-27
    assert moll(almost, 1.0)[0] == pytest.approx(0, abs=1e-9)
    for edges in (perp, almost):
        a, p = _p(edges)
        assert cd.ipco_unit_mollifier_threshold(p) == pytest.approx(0.016)  # :255-260
    eps_x = cd.ipco_unit_mollifier_threshold(_p(rest)[1])
    for edges in (perp, almost):
        s, m, g, H = moll(edges, eps_x)
        assert 0 <= m <= 1
        x = edges.ravel()
        fg = _fd_grad(lambda y: moll(y.reshape(4, 3), eps_x)[1], x)
        assert np.allclose(g, fg, rtol=1e-5, atol=1e-6 * max(1, np.abs(fg).max()))
        fH = np.array([_fd_grad(lambda y, i=i: moll(y.reshape(4, 3), eps_x)[2][i], x) for i in range(12)])
        assert np.allclose(H, fH, rtol=1e-4, atol=1e-5 * max(1, np.abs(fH).max()))
    # scalar mollifier (:82-160) through a perpendicular pair scaled so that its cross norm equals x
    for rel_x in (0, 0.5, 1, 2):
        for eps in (1e-3, 1e-1, 1, 2):
            x = rel_x * eps
            L = np.sqrt(np.sqrt(x)) if x > 0 else 0.0  # |u x v|^2 = L^4 for perpendicular edges of length L
            e = np.array([[0.0, 0, 0], [L, 0, 0], [0, 0, 1], [0, L, 1]])
            s, m, _, _ = moll(e, eps)
            assert s == pytest.approx(x, rel=1e-12, abs=1e-300)
            assert 0 <= m <= 1
            if x > eps:
                assert m == 1
            else:
                assert m == pytest.approx((-x / eps + 2) * (x / eps))


def test_barrier_derivatives(oracle):
    """barrier/test_barrier.cpp:398-462 (ClampedLogBarrier section) + :11-43 of barrier.cpp"""
    cd = oracle.cdll

    def b(d, dhat):
        out, op = _p(np.zeros(3))
        cd.ipco_unit_barrier(d, dhat, op)
        return out

    for use_sqr in (False, True):
        for e in range(-2 if use_sqr else -5, 0):
            dhat = 10.0 ** e
            for d in np.arange(0.5 * dhat, 0.9 * dhat, (0.9 - 0.5) / 10.0 * dhat)[:10]:
                dd, hh = (d * d, dhat * dhat) if use_sqr else (d, dhat)
                h = 1e-7 * dd
                f0, f1, f2 = b(dd, hh)
                assert f1 == pytest.approx((b(dd + h, hh)[0] - b(dd - h, hh)[0]) / (2 * h), rel=1e-5)
                assert f2 == pytest.approx((b(dd + h, hh)[1] - b(dd - h, hh)[1]) / (2 * h), rel=1e-5)
    assert b(1.0, 1.0)[0] == 0 and b(2.0, 1.0).tolist() == [0, 0, 0] and np.isinf(b(0.0, 1.0)[0])


def test_project_to_psd(oracle):
    """utils/test_utils.cpp:27-48"""
    cd = oracle.cdll

    def psd(A, mode):
        a, p = _p(np.array(A, float).T.copy())
        assert cd.ipco_unit_project_to_psd(a.shape[0], p, mode) == 0
        return a.T

    I3 = np.eye(3)
    assert np.allclose(psd(I3, 1), I3) and np.allclose(psd(I3, 2), I3)
    assert np.allclose(psd(-I3, 1), 0)
    A = [[2, 1], [1, 2]]
    assert np.allclose(psd(A, 1), A) and np.allclose(psd(A, 2), A)
    assert np.allclose(psd([[1, 2], [2, 1]], 1), 1.5 * np.ones((2, 2)))  # eigenvalues -1, 3
    assert np.allclose(psd([[1, 2], [2, 1]], 2), [[2, 1], [1, 2]])
    rng = np.random.default_rng(0)
    for n in (6, 9, 12):
        M = rng.normal(size=(n, n))
        M = M + M.T
        w, V = np.linalg.eigh(M)
        assert np.allclose(psd(M, 1), (V * np.maximum(w, 0)) @ V.T, atol=1e-12)
        assert np.allclose(psd(M, 2), (V * np.abs(w)) @ V.T, atol=1e-12)


def test_morton_code(oracle):
    """math/morton.hpp:23-63 compiled standalone from the reference tree gives 0x5600000000000000"""
    assert oracle.cdll.ipco_unit_morton_3D(0.5, 0.25, 0.75) == 0x5600000000000000


# ---- broad phase, collisions, potential, step size (shared with the GPU tests) --------------------------------------
def check_broad_phase_kats(api):
    """broad_phase/test_broad_phase.cpp:116-137 (2D embedded at z = 0), :207-231/:246-253 (crossing edges), :273-294
    (100 chained boxes -> 99 vertex-vertex candidates)"""
    V0 = np.array([[1.11111, 0.5, 0], [1.11111, 0.75, 0], [1, 0.5, 0], [1, 0.75, 0]])
    V1 = V0.copy()
    V1[:2, 0] = 0.888889
    E = np.array([[1, 0], [2, 3]])
    mesh = api.CollisionMesh(V0, E)
    cand = api.Candidates()
    cand.build(mesh, V0, V1, 0.0)
    assert len(cand.ee_candidates) == 1  # the two edges sweep through each other
    assert not cand.is_step_collision_free(mesh, V0, V1)

    V0 = np.array([[-1.0, -1, 0], [1, -1, 0], [0, 1, 1], [0, 1, -1]])
    U = np.zeros_like(V0)
    U[:2, 1], U[2:, 1] = 2, -2
    E = np.array([[0, 1], [2, 3]])
    mesh = api.CollisionMesh(V0, E)
    cand = api.Candidates()
    cand.build(mesh, V0, V0 + U, 0.0)
    assert [tuple(c) for c in cand.ee_candidates] == [(0, 1)]
    toi = cand.compute_collision_free_stepsize(mesh, V0, V0 + U)
    assert 0.4 < toi <= 0.5  # the edges meet at t = 0.5

    Vc = np.zeros((100, 3))
    Vc[:, 0] = 0.6 * np.arange(100) + 0.5
    mesh = api.CollisionMesh(Vc)
    cand = api.Candidates()
    cand.build(mesh, Vc, None, 0.5)
    assert len(cand.vv_candidates) == 99 and cand.size() == 99


def check_codim_kats(api):
    """collisions/test_normal_collisions.cpp:14-107 and :109-209"""
    V = np.array([[0, 0, 0], [0, 0, 1], [0, 1, 0], [0, 1, 1], [1, 0, 0], [1, 0, 1], [1, 1, 0], [1, 1, 1]], float)
    V -= V.mean(0)
    mesh = api.CollisionMesh(V)
    assert (mesh.num_vertices(), mesh.num_codim_vertices(), mesh.num_edges(), mesh.num_faces()) == (8, 8, 0, 0)
    for area in (False, True):
        for physical in (False, True):
            c = api.NormalCollisions()
            c.set_use_area_weighting(area)
            c.build(mesh, V, 0.25, 0.8)
            assert c.counts() == [12, 0, 0, 0]
            B = api.BarrierPotential(0.25, 1.0, physical)
            assert B(c, mesh, V) > 0
            f = -B.gradient(c, mesh, V).reshape(-1, 3)
            assert np.allclose(f / np.linalg.norm(f, axis=1, keepdims=True), V / np.linalg.norm(V, axis=1, keepdims=True))
    V1 = V.copy()
    V1[:, 1] *= 0.5
    cand = api.Candidates()
    cand.build(mesh, V, V1, 0.4)
    assert cand.size() == len(cand.vv_candidates) > 0
    assert not cand.is_step_collision_free(mesh, V, V1, 0.8)
    assert cand.compute_collision_free_stepsize(mesh, V, V1, 0.8) == pytest.approx((1 - (0.8 + 1e-4)) / 2 / 0.25, rel=1.2e-5)

    V = np.array([[0, 0, 0], [1, 0, 0], [0, 0, -1], [-1, 0, 0], [0, 0, 1], [0, 1, 0], [0, 2, 0], [0, 3, 0]], float)
    E = np.array([[0, 1], [0, 2], [0, 3], [0, 4]])
    mesh = api.CollisionMesh(V, E)
    assert (mesh.num_codim_vertices(), mesh.num_codim_edges(), mesh.num_edges(), mesh.num_faces()) == (3, 4, 4, 0)
    V1 = V.copy()
    V1[5:, 1] -= 4
    cand = api.Candidates()
    cand.build(mesh, V, V1, 1e-3)
    assert [len(cand.vv_candidates), len(cand.ev_candidates), len(cand.ee_candidates), len(cand.fv_candidates)] == [3, 12, 0, 0]
    assert not cand.is_step_collision_free(mesh, V, V1, 2e-3)
    assert cand.compute_collision_free_stepsize(mesh, V, V1, 2e-3) == pytest.approx((1 - (2e-3 + 1e-4)) / 4, rel=1.2e-5)
    for area in (False, True):
        c = api.NormalCollisions()
        c.set_use_area_weighting(area)
        c.build(mesh, V, 0.25, 0.8)
        assert c.counts() == [2, 4, 0, 0]
        assert api.BarrierPotential(0.25, 1.0)(c, mesh, V) > 0


def check_barrier_potential_scenes(api):
    """potential/test_barrier_potential.cpp:133-245: collisions exist and the gradient / Hessian match finite differences
    of the potential on the SAME collision set"""
    for name, (V, E, F, dhat) in kats.barrier_potential_scenes().items():
        if E is None:
            E = api_edges(F)
        mesh = api.CollisionMesh(V, E, F)
        for area in (False, True):
            for physical in (False, True):
                c = api.NormalCollisions()
                c.set_use_area_weighting(area)
                c.build(mesh, V, dhat)
                assert not c.empty(), name
                B = api.BarrierPotential(dhat, 1.0, physical)
                g = B.gradient(c, mesh, V)
                H = B.hessian(c, mesh, V).toarray()
                x = V.ravel().copy()
                h = 1e-9
                fg = np.zeros_like(x)
                fH = np.zeros((x.size, x.size))
                for i in range(x.size):
                    e = np.zeros_like(x)
                    e[i] = h
                    fg[i] = (B(c, mesh, (x + e).reshape(V.shape)) - B(c, mesh, (x - e).reshape(V.shape))) / (2 * h)
                    fH[:, i] = (B.gradient(c, mesh, (x + e).reshape(V.shape)) - B.gradient(c, mesh, (x - e).reshape(V.shape))) / (2 * h)
                assert np.allclose(g, fg, rtol=2e-4, atol=1e-6 * np.abs(fg).max()), name
                assert np.allclose(H, fH, rtol=2e-3, atol=1e-5 * np.abs(fH).max()), name
                assert np.allclose(H, H.T, atol=1e-12 * np.abs(H).max())
                Hp = B.hessian(c, mesh, V, 1).toarray()
                assert np.linalg.eigvalsh(0.5 * (Hp + Hp.T)).min() >= -1e-9 * np.abs(Hp).max()


def api_edges(F):
    e = np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]]).astype(np.int64)
    e.sort(axis=1)
    return np.unique(e, axis=0).astype(np.int32)


def check_readme_quick_start(api):
    """README.md:51-88"""
    V, E, F, V1, dhat = kats.readme_quick_start()
    mesh = api.CollisionMesh(V, E, F)
    c = api.NormalCollisions()
    c.build(mesh, V, dhat)
    assert c.size() > 0  # "the two triangles are within dhat, so collisions are active"
    B = api.BarrierPotential(dhat, 1e3)
    assert B(c, mesh, V) > 0
    assert np.abs(B.gradient(c, mesh, V)).max() > 0
    H = B.hessian(c, mesh, V)
    assert H.shape == (18, 18) and H.nnz > 0
    step = api.compute_collision_free_stepsize(mesh, V, V1)
    assert 0.3 < step <= 0.5  # triangle 2 reaches triangle 1 at t = 0.5


def check_ccd_kats(api):
    """ccd/test_edge_edge_ccd.cpp, test_point_triangle_ccd.cpp, test_point_edge_ccd.cpp, test_point_point_ccd.cpp"""
    for kind, cases in ((2, kats.edge_edge_ccd_cases()), (3, kats.point_triangle_ccd_cases())):
        groups = {}
        for c in cases:
            groups.setdefault((c["tol"], c["max_iter"], c["tmax"]), []).append(c)
        for (tol, max_iter, tmax), cs in groups.items():
            t0, t1 = np.stack([c["t0"] for c in cs]), np.stack([c["t1"] for c in cs])
            for ccd in (api.TightInclusionCCD(tol, max_iter), api.AdditiveCCD()):
                hit, toi = api.narrow_phase_ccd(kind, t0, t1, 0.0, tmax, ccd)
                for c, h, t in zip(cs, hit, toi):
                    if c["conservative"]:
                        assert h or not c["expected"], (c["name"], type(ccd).__name__)
                    else:
                        assert h == c["expected"], (c["name"], type(ccd).__name__, c["tol"])
                    if h:
                        assert 0 <= t <= tmax
    pe = kats.point_edge_ccd_cases()
    t0, t1 = np.stack([c["t0"] for c in pe]), np.stack([c["t1"] for c in pe])
    for ccd in (api.TightInclusionCCD(), api.AdditiveCCD(-1, 0.999)):
        hit, toi = api.narrow_phase_ccd(1, t0, t1, 0.0, 1.0, ccd)
        for c, h, t in zip(pe, hit, toi):
            assert h == c["expected"], (c["name"], type(ccd).__name__)
            if h:
                # the reference asserts toi <= expected (conservative); the lower bound only guards against a trivial 0
                assert t <= c["toi"] + 1e-9 and t >= 0.9 * c["toi"], (c["name"], t)
    pp = kats.point_point_ccd_cases()
    for md in pp["min_distances"]:
        for ccd in (api.TightInclusionCCD(), api.AdditiveCCD(-1, 0.999)):
            hit, toi = api.narrow_phase_ccd(0, pp["t0"][None], pp["t1"][None], md, 1.0, ccd)
            assert hit[0] and toi[0] == pytest.approx(pp["toi"], abs=pp["margin"] + md)


def test_broad_phase_kats(oracle):
    check_broad_phase_kats(oracle)


def check_faces_to_edges_kats(api):
    """tests/src/tests/test_collision_mesh.cpp:70-110 (CollisionMesh::construct_faces_to_edges, collision_mesh.cpp:510-543)"""
    V = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], float)
    F = np.array([[0, 1, 2]], np.int32)
    for E, want in (([[0, 1], [1, 2], [2, 0]], [0, 1, 2]), ([[2, 0], [2, 1], [1, 0]], [2, 1, 0]), ([[0, 1], [2, 0], [2, 1]], [0, 2, 1])):
        mesh = api.CollisionMesh(V, np.array(E, np.int32), F)
        assert mesh.faces_to_edges().tolist() == [want]
    with pytest.raises(RuntimeError, match="Unable to find edge!"):
        api.CollisionMesh(V, np.array([[0, 1], [1, 2], [0, 3]], np.int32), F)
    # codim points (:112-122): every vertex of a mesh without edges is codimensional
    assert api.CollisionMesh(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0]], float)).num_codim_vertices() == 4


def test_faces_to_edges_kats(oracle):
    check_faces_to_edges_kats(oracle)


def test_codim_kats(oracle):
    check_codim_kats(oracle)


def test_barrier_potential_scenes(oracle):
    check_barrier_potential_scenes(oracle)


def test_readme_quick_start(oracle):
    check_readme_quick_start(oracle)


def test_ccd_kats(oracle):
    check_ccd_kats(oracle)


# ---- CollisionSetType::IMPROVED_MAX_APPROX: restated in the oracle only so far ----------------------
def _fan(k=6, apex=0.3):
    """a cone of k triangles around the raised vertex 0 (interior: every spoke is shared by two triangles; the apex and
    the spokes are convex, so their closest-point regions have a finite size and no test point sits on a type
    boundary) + one free point"""
    ang = 2 * np.pi * np.arange(k) / k
    V = np.vstack([[0, 0, apex], np.c_[np.cos(ang), np.sin(ang), np.zeros(k)], [0, 0, 0]])
    F = np.array([[0, 1 + i, 1 + (i + 1) % k] for i in range(k)], np.int32)
    return V, F


def test_improved_max_approx_counts_and_weights(oracle):
    check_improved_max_approx_kats(oracle)


def check_improved_max_approx_kats(oracle):
    """(the parameter is any API namespace: the oracle here, the CUDA library in the GPU tests)
    collisions/test_normal_collisions.cpp:69-107,170-209 (12 vertex-vertex collisions on the cube for every set type;
    6 + 1 collisions, 2 + 1 of them vertex-vertex, on the edge-vertex scene) and the defining property of the
    convergent formulation (normal_collisions_builder.cpp:340-543): a point above an interior VERTEX of valence k is
    counted k times by the IPC set (once per incident face) and once after the corrections (k - k + 1); above the
    middle of an interior EDGE twice by IPC and once after the corrections (2 - 1)."""
    T = oracle.NormalCollisions.CollisionSetType

    def build(mesh, V, dhat, dmin, t, area=False):
        c = oracle.NormalCollisions()
        c.set_use_area_weighting(area)
        c.set_collision_set_type(t)
        c.build(mesh, V, dhat, dmin)
        return c

    V = np.array([[0, 0, 0], [0, 0, 1], [0, 1, 0], [0, 1, 1], [1, 0, 0], [1, 0, 1], [1, 1, 0], [1, 1, 1]], float)
    V -= V.mean(0)
    mesh = oracle.CollisionMesh(V)
    for area in (False, True):
        assert build(mesh, V, 0.25, 0.8, T.IMPROVED_MAX_APPROX, area).counts() == [12, 0, 0, 0]

    V = np.array([[0, 0, 0], [1, 0, 0], [0, 0, -1], [-1, 0, 0], [0, 0, 1], [0, 1, 0], [0, 2, 0], [0, 3, 0]], float)
    E = np.array([[0, 1], [0, 2], [0, 3], [0, 4]])
    mesh = oracle.CollisionMesh(V, E)
    for area in (False, True):
        c = build(mesh, V, 0.25, 0.8, T.IMPROVED_MAX_APPROX, area)
        assert c.counts() == [3, 4, 0, 0]
        assert oracle.BarrierPotential(0.25, 1.0)(c, mesh, V) > 0
        extra = dict(zip(map(tuple, c.vv_collisions.ids), c.vv_collisions.weight))[(0, 5)]
        assert extra == (1 - 4) * (0.5 * mesh.vertex_areas()[5] if area else 1.0)  # 4 incident edges at vertex 0

    k, dhat = 6, 0.1
    V, F = _fan(k)
    E = oracle.edges_from_faces(F)
    p = k + 1
    mesh = oracle.CollisionMesh(V, E, F)
    # above the interior vertex
    B = oracle.BarrierPotential(dhat, 1.0)

    def summary(X, t):  # (counts, records, energy) of one build; a handle is valid until the next build on its mesh
        c = build(mesh, X, dhat, 0.0, t)
        recs = {n: (getattr(c, n + "_collisions").ids.tolist(), getattr(c, n + "_collisions").weight.tolist()) for n in ("vv", "ev", "ee", "fv")}
        return c.counts(), recs, B(c, mesh, X)

    def normal(f):
        n = np.cross(V[f[1]] - V[f[0]], V[f[2]] - V[f[0]])
        return n / np.linalg.norm(n)

    X = V.copy()
    X[p] = V[0] + [0.004 * dhat, 0.002 * dhat, 0.5 * dhat]  # in the apex's region, slightly off the axis
    (n_ipc, r_ipc, e_ipc), (n_imp, r_imp, e_imp) = summary(X, T.IPC), summary(X, T.IMPROVED_MAX_APPROX)
    assert n_ipc == [1, 0, 0, 0] and r_ipc["vv"] == ([[0, p]], [float(k)])
    assert n_imp == [1, 0, 0, 0] and r_imp["vv"] == ([[0, p]], [1.0])
    assert e_ipc == pytest.approx(k * e_imp, rel=1e-14)
    # above the middle of an interior edge (spoke 0-1)
    ridge = normal(F[0]) + normal(F[k - 1])  # spoke 0-1 is shared by the first and the last triangle
    X[p] = 0.5 * (V[0] + V[1]) + 0.5 * dhat * ridge / np.linalg.norm(ridge) + [0, 0.003 * dhat, 0]
    (n_ipc, r_ipc, _), (n_imp, r_imp, _) = summary(X, T.IPC), summary(X, T.IMPROVED_MAX_APPROX)
    assert n_ipc == [0, 1, 0, 0] and r_ipc["ev"][1] == [2.0]
    assert n_imp == [0, 1, 0, 0] and r_imp["ev"][1] == [1.0] and r_imp["ev"][0] == r_ipc["ev"][0]
    # above the inside of a face: nothing to correct
    X[p] = V[F[0]].mean(0) + 0.5 * dhat * normal(F[0])
    assert summary(X, T.IPC)[0] == summary(X, T.IMPROVED_MAX_APPROX)[0] == [0, 0, 0, 1]


def check_improved_max_approx_builders(api, scenes, world=3):
    """IMPROVED_MAX_APPROX built by several builders over disjoint candidate shards (IPCB_DEFER_CORRECTIONS: deferred build,
    exchange of the sub-element pairs, every builder adds the corrections of its slice) unites to the single-builder set"""
    import types

    for name, area in (("stack", False), ("stack", True), ("drape", False), ("dense", True)):
        V0, V1, E, F, P = {"stack": lambda: scenes.cloth_stack(3, 20), "drape": lambda: scenes.cloth_on_sphere(32, 16, drape=True),
                           "dense": lambda: scenes.dense_sheet(8, 2.0)}[name]()
        dhat = P["dhat"]
        IMA = api.NormalCollisions.CollisionSetType.IMPROVED_MAX_APPROX
        mesh = api.CollisionMesh(V0, E, F)
        cand = api.Candidates()
        cand.build(mesh, V0, 0.5 * dhat)
        one = api.NormalCollisions()
        one.set_use_area_weighting(area), one.set_collision_set_type(IMA)
        one.build(cand, mesh, V0, dhat)
        want = [getattr(one, k + "_collisions") for k in ("vv", "ev", "ee", "fv")]
        full = [np.asarray(getattr(cand, k + "_candidates")).copy() for k in ("vv", "ev", "ee", "fv")]
        meshes = [api.CollisionMesh(V0, E, F) for _ in range(world)]
        builders = []
        for r, m in enumerate(meshes):
            c = api.Candidates()
            c.set(m, *[a[(len(a) * r) // world:(len(a) * (r + 1)) // world] for a in full])
            b = api.NormalCollisions()
            b.set_use_area_weighting(area), b.set_collision_set_type(IMA)
            b.build(c, m, V0, dhat, defer_corrections=True)
            builders.append(b)
        keys = [b.correction_keys() for b in builders]
        assert sum(len(k) for ks in keys for k in ks) > 0
        united = [np.concatenate([ks[k] for ks in keys]) for k in range(4)]
        parts = []
        for r, b in enumerate(builders):
            b.apply_corrections(united, r, world)
            parts.append([getattr(b, k + "_collisions") for k in ("vv", "ev", "ee", "fv")])
        m = api.NormalCollisions()
        m.assign(mesh, [[types.SimpleNamespace(ids=x.ids, weight=x.weight, eps_x=x.eps_x, dtype=x.dtype) for x in p] for p in parts], 0.0)
        have = [getattr(m, k + "_collisions") for k in ("vv", "ev", "ee", "fv")]
        for kind, (sa, sb) in enumerate(zip(have, want)):
            scale = max(np.abs(sb.weight).max(), 1e-300) if len(sb.weight) else 1.0
            ka, kb = np.abs(sa.weight) > 1e-12 * scale, np.abs(sb.weight) > 1e-12 * scale
            assert np.array_equal(sa.ids[ka], sb.ids[kb]), (name, area, kind)
            assert np.array_equal(sa.dtype[ka], sb.dtype[kb]) and np.array_equal(sa.eps_x[ka], sb.eps_x[kb])
            assert np.allclose(sa.weight[ka], sb.weight[kb], rtol=1e-12, atol=0)


def test_improved_max_approx_over_several_builders(oracle, scenes):
    check_improved_max_approx_builders(oracle, scenes)


def test_improved_max_approx_derivatives(oracle, scenes):
    check_improved_max_approx_derivatives(oracle, scenes)


def check_improved_max_approx_derivatives(oracle, scenes):
    """potential/test_barrier_potential.cpp:34,126 run their finite-difference checks for IMPROVED_MAX_APPROX too: negative
    weights and mollified edge-edge collisions with vertex / edge distance types go through the same potential"""
    T = oracle.NormalCollisions.CollisionSetType
    V0, V1, E, F, P = scenes.dense_sheet(6, 1.5)  # dhat spans 1.5 cells: vertex / edge proximity everywhere
    dhat = P["dhat"]
    mesh = oracle.CollisionMesh(V0, E, F)
    for area in (False, True):
        ipc = oracle.NormalCollisions()
        ipc.set_use_area_weighting(area)
        ipc.build(mesh, V0, dhat)
        n_ipc = ipc.counts()
        c = oracle.NormalCollisions()
        c.set_use_area_weighting(area)
        c.set_collision_set_type(T.IMPROVED_MAX_APPROX)
        c.build(mesh, V0, dhat)
        assert sum(c.counts()) > 0 and c.counts() != n_ipc  # corrections were added
        assert (c.ee_collisions.dtype != 8).any()  # mollified edge-edge corrections with vertex / edge distance types
        w = np.concatenate([getattr(c, k + "_collisions").weight for k in ("vv", "ev", "ee", "fv")])
        assert (w < 0).any() and (w > 0).any()
        B = oracle.BarrierPotential(dhat, 1.0, use_physical_barrier=area)
        g = B.gradient(c, mesh, V0)
        H = B.hessian(c, mesh, V0)
        rng = np.random.default_rng(5)
        p = rng.standard_normal(V0.shape)
        eps = 1e-6 * dhat
        fd_e = (B(c, mesh, V0 + eps * p) - B(c, mesh, V0 - eps * p)) / (2 * eps)
        fd_g = (B.gradient(c, mesh, V0 + eps * p) - B.gradient(c, mesh, V0 - eps * p)) / (2 * eps)
        assert fd_e == pytest.approx(g @ p.ravel(), rel=1e-5)
        assert np.linalg.norm(fd_g - H @ p.ravel()) <= 1e-5 * np.linalg.norm(fd_g)
        Hp = B.hessian(c, mesh, V0, oracle.PSDProjectionMethod.CLAMP).toarray()
        assert np.linalg.eigvalsh(0.5 * (Hp + Hp.T)).min() >= -1e-9 * np.abs(Hp).max()


def check_independent_collision_sets(api, scenes):
    """the reference's NormalCollisions are plain containers: any number of them may exist per mesh (a lagged set for
    friction, the sets of a line search).  Here each object owns a collision-set object of the library and is swapped in
    when it is used (ipcb_collisions_swap)."""
    V0, V1, E, F, P = scenes.cloth_stack(3, 10)
    dhat = P["dhat"]
    mesh = api.CollisionMesh(V0, E, F)
    B = api.BarrierPotential(dhat, 1.0)
    a = api.NormalCollisions()
    a.build(mesh, V0, dhat)
    ea, ga, ids_a = B(a, mesh, V0), B.gradient(a, mesh, V0), a.ee_collisions.ids.copy()
    X = V0 + 0.3 * (V1 - V0)
    b = api.NormalCollisions()
    b.build(mesh, X, dhat)  # a second set on the same mesh: `a` stays valid
    eb = B(b, mesh, X)
    assert b.counts() != a.counts() or not np.array_equal(b.ee_collisions.ids, ids_a)
    close = lambda x, y: abs(x - y) <= 1e-13 * abs(y)  # the oracle's parallel reduction order is not fixed
    assert close(B(a, mesh, V0), ea) and np.allclose(B.gradient(a, mesh, V0), ga, rtol=1e-12, atol=0)  # `a` swapped back in
    assert np.array_equal(a.ee_collisions.ids, ids_a)
    assert close(B(b, mesh, X), eb)  # and `b` again
    Ha, Hb = B.hessian(a, mesh, V0), B.hessian(b, mesh, X)
    assert Ha.nnz > 0 and Hb.nnz > 0 and (Ha.nnz != Hb.nnz or abs(Ha - Hb).max() > 0)
    a.build(mesh, X, dhat)  # rebuilding one set leaves the other alone
    assert a.counts() == b.counts() and close(B(a, mesh, X), eb) and close(B(b, mesh, X), eb)
    with pytest.raises(RuntimeError):
        B(api.NormalCollisions(), mesh, V0)  # never built


def test_independent_collision_sets_oracle(oracle):
    import ipctk_b200

    check_independent_collision_sets(oracle, ipctk_b200._pkg.scenes)
