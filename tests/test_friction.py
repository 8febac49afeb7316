"""Friction: TangentialCollisions::build + FrictionPotential energy / gradient / Hessian.

Reference tests restated (they are finite-difference based and generate their data, so they port without fixtures):
tests/src/tests/friction/test_smooth_friction_mollifier.cpp:11-95 and test_smooth_mu.cpp (derivative identities of the
mollifier family), tests/src/tests/potential/test_friction_potential.cpp:16-58 (gradient and Hessian of the potential
against finite differences of the energy / gradient), plus the defining properties of the lagged quantities (orthonormal
tangent basis orthogonal to the contact normal, closest points, N = -kappa b' 2 d).  CPU suite: the oracle; GPU suite: the
CUDA path against the oracle (records bit-identical in ids, values to 1e-10) and the same FD checks."""
import ctypes as C

import numpy as np
import pytest


def _scene(scenes, name):
    return {"stack": lambda: scenes.cloth_stack(3, 10), "drape": lambda: scenes.cloth_on_sphere(16, 10, drape=True),
            "soup": lambda: scenes.random_soup(60, seed=7)}[name]()


def build_sets(api, V0, E, F, dhat, mu_s=0.5, mu_k=0.3, kappa=1e3):
    mesh = api.CollisionMesh(V0, E, F)
    c = api.NormalCollisions()
    c.build(mesh, V0, dhat)
    B = api.BarrierPotential(dhat, kappa)
    t = api.TangentialCollisions()
    nV = V0.shape[0]
    rng = np.random.default_rng(1)
    t.build(mesh, V0, c, B, mu_s * (1 + 0.2 * rng.random(nV)), mu_k * (1 + 0.2 * rng.random(nV)))
    return mesh, c, t


def check_lagged_quantities(api, scenes, name):
    V0, V1, E, F, P = _scene(scenes, name)
    dhat = P["dhat"]
    mesh, c, t = build_sets(api, V0, E, F, dhat)
    assert t.counts()[0] == c.counts()[0] and t.counts()[1] == c.counts()[1] and t.counts()[3] == c.counts()[3]
    assert t.counts()[2] <= c.counts()[2] and t.size() > 0
    for kind, key in enumerate(("vv", "ev", "ee", "fv")):
        r = getattr(t, key + "_collisions")
        if not len(r.ids):
            continue
        Pm = r.tangent_basis  # n x 2 x 3
        gram = np.einsum("nik,njk->nij", Pm, Pm)
        assert np.allclose(gram, np.eye(2)[None], atol=1e-12), key  # orthonormal columns
        assert np.all(r.normal_force_magnitude > 0) and np.all(r.weight > 0) and np.all(r.mu_s > r.mu_k)
        # the contact normal is orthogonal to the basis, the closest points realise the distance
        if key == "fv":
            f = F[r.ids[:, 0]]
            p, t0, t1, t2 = V0[r.ids[:, 1]], V0[f[:, 0]], V0[f[:, 1]], V0[f[:, 2]]
            q = t0 + r.closest_point[:, :1] * (t1 - t0) + r.closest_point[:, 1:] * (t2 - t0)
            nrm = p - q
            assert np.abs(np.einsum("nik,nk->ni", Pm, nrm)).max() <= 1e-9 * np.abs(nrm).max()
        if key == "ee":
            ea, eb = E[r.ids[:, 0]], E[r.ids[:, 1]]
            pa = V0[ea[:, 0]] + r.closest_point[:, :1] * (V0[ea[:, 1]] - V0[ea[:, 0]])
            pb = V0[eb[:, 0]] + r.closest_point[:, 1:] * (V0[eb[:, 1]] - V0[eb[:, 0]])
            nrm = pa - pb
            assert np.abs(np.einsum("nik,nk->ni", Pm, nrm)).max() <= 1e-8 * np.abs(nrm).max()
            d = np.linalg.norm(nrm, axis=1)
            # N = -kappa b'(d^2) 2 d with the clamped log barrier b(x) = -(x - xhat)^2 ln(x / xhat), xhat = dhat^2
            x, xh = d * d, dhat * dhat
            db = -(2 * (x - xh) * np.log(x / xh) + (x - xh) ** 2 / x)
            assert np.allclose(r.normal_force_magnitude, -1e3 * db * 2 * d, rtol=1e-9)
    return t


def check_potential_fd(api, scenes, name, eps_v):
    """test_friction_potential.cpp:16-58: gradient / Hessian against central differences"""
    V0, V1, E, F, P = _scene(scenes, name)
    mesh, c, t = build_sets(api, V0, E, F, P["dhat"])
    rng = np.random.default_rng(3)
    U = rng.normal(0, 1.0, V0.shape) * eps_v * np.array([3.0, 0.3, 1.0])[rng.integers(0, 3, V0.shape[0])][:, None]  # slow, fast and mixed contacts
    D = api.FrictionPotential(eps_v)
    e, g = D(t, mesh, U), D.gradient(t, mesh, U)
    H = D.hessian(t, mesh, U)
    assert e > 0 and H.nnz > 0 and abs(H - H.T).max() <= 1e-10 * abs(H).max()
    p = rng.standard_normal(V0.shape)
    h = 1e-6 * eps_v
    fd_e = (D(t, mesh, U + h * p) - D(t, mesh, U - h * p)) / (2 * h)
    assert abs(fd_e - g @ p.ravel()) <= 1e-6 * abs(g @ p.ravel())
    fd_g = (D.gradient(t, mesh, U + h * p) - D.gradient(t, mesh, U - h * p)) / (2 * h)
    Hp = H @ p.ravel()
    assert np.linalg.norm(fd_g - Hp) <= 1e-5 * np.linalg.norm(Hp)
    for mode in (api.PSDProjectionMethod.CLAMP, api.PSDProjectionMethod.ABS):  # projected: PSD along random directions
        Hm = D.hessian(t, mesh, U, mode)
        for _ in range(3):
            v = rng.standard_normal(3 * V0.shape[0])
            assert v @ (Hm @ v) >= -1e-9 * abs(Hm).max() * (v @ v)
    # at rest (|u| = 0) and fully sliding (|u| > eps_v everywhere) branches
    Z = np.zeros_like(U)
    assert np.abs(D.gradient(t, mesh, Z)).max() == 0.0 and D.hessian(t, mesh, Z).nnz > 0
    S = U * 1e3
    Hs = D.hessian(t, mesh, S)
    v = rng.standard_normal(3 * V0.shape[0])
    assert v @ (Hs @ v) >= -1e-9 * abs(Hs).max() * (v @ v)  # mu N f1/|u| (I - u u^T/|u|^2) is PSD without projection
    return dict(e=e, g=g, H=H)


def test_mollifier_identities(oracle):
    """test_smooth_friction_mollifier.cpp / test_smooth_mu.cpp: f1 = f0', f2 = f1', f1/x and (f2 x - f1)/x^3 consistent —
    observed through a one-contact potential: E(u) = w N mu_f0(|u|), so dE/d|u| = w N mu_f1 and so on"""
    cd = oracle.cdll
    for fn in ("ipco_unit_smooth_mu_f0", "ipco_unit_smooth_mu_f1_over_x", "ipco_unit_smooth_mu_f2_x_minus_mu_f1_over_x3"):
        getattr(cd, fn).restype = C.c_double
        getattr(cd, fn).argtypes = [C.c_double] * 4
    for eps_v in (1e-3, 1e-1):
        for mu_s, mu_k in ((0.5, 0.5), (0.6, 0.3)):
            for x in np.array([0.05, 0.3, 0.49, 0.51, 0.8, 0.999, 1.5, 10.0]) * eps_v:
                h = 1e-6 * x
                f0 = lambda y: cd.ipco_unit_smooth_mu_f0(y, mu_s, mu_k, eps_v)
                f1ox = lambda y: cd.ipco_unit_smooth_mu_f1_over_x(y, mu_s, mu_k, eps_v)
                f1 = (f0(x + h) - f0(x - h)) / (2 * h)
                assert f1ox(x) * x == pytest.approx(f1, rel=1e-5)
                f2 = (f1ox(x + h) * (x + h) - f1ox(x - h) * (x - h)) / (2 * h)
                g = cd.ipco_unit_smooth_mu_f2_x_minus_mu_f1_over_x3(x, mu_s, mu_k, eps_v)
                assert g * x ** 3 == pytest.approx(f2 * x - f1ox(x) * x, rel=2e-4, abs=1e-7)
            assert f0(0.0) > 0 and f0(eps_v) == pytest.approx(mu_k * eps_v, rel=1e-12)  # C^1 match with mu_k |u| at eps_v


@pytest.mark.parametrize("name", ["stack", "drape", "soup"])
def test_lagged_quantities_oracle(oracle, scenes, name):
    check_lagged_quantities(oracle, scenes, name)


@pytest.mark.parametrize("name", ["stack", "soup"])
def test_friction_potential_fd_oracle(oracle, scenes, name):
    check_potential_fd(oracle, scenes, name, 1e-3)


def check_reference_tangent_kats(api):
    """the reference's known answers for tangent bases and closest points (tests/src/tests/tangent/test_tangent_basis.cpp:12-31,
    51-70,86-105,119-137; test_closest_point.cpp:48-56,78-83), reached through TangentialCollisions::build on one-collision meshes"""
    X, Y, Z = np.eye(3)

    def one(V, E, F, key, dhat=2.0):
        V = np.asarray(V, float)
        E = np.asarray(E, np.int32).reshape(-1, 2)
        F = np.asarray(F, np.int32).reshape(-1, 3)
        mesh = api.CollisionMesh(V, E, F)
        c = api.NormalCollisions()
        c.build(mesh, V, dhat)
        counts = dict(zip(("vv", "ev", "ee", "fv"), c.counts()))
        assert counts[key] == 1 and sum(counts.values()) == 1, counts
        t = api.TangentialCollisions()
        t.build(mesh, V, c, api.BarrierPotential(dhat, 1.0), np.full(len(V), 0.5), np.full(len(V), 0.5))
        r = getattr(t, key + "_collisions")
        assert len(r.ids) == 1
        return r.tangent_basis[0], r.closest_point[0]  # 2 x 3 (the columns of the 3 x 2 basis), closest-point coordinates

    # point-triangle: p above the triangle's plane y = 0 (the reference evaluates the basis for exactly these points)
    B, cp = one([[0, 1, 0], [-1, 0, 1], [1, 0, 1], [0, 0, -1]], [[1, 2], [2, 3], [3, 1]], [[1, 2, 3]], "fv")
    assert abs(abs(B[0] @ X) - 1) < 1e-12 and abs(abs(B[1] @ Z) - 1) < 1e-12
    assert np.allclose(cp, [0.25, 0.5], atol=1e-12)  # t0 + u (t1 - t0) + v (t2 - t0) = (0, 0, 0)
    # edge-edge: the reference's crossing edges, the second one lifted by 1 (the basis does not depend on the lift)
    B, cp = one([[-1, 0, 0], [1, 0, 0], [0, 1, -1], [0, 1, 1]], [[0, 1], [2, 3]], [], "ee")
    assert abs(abs(B[0] @ X) - 1) < 1e-12 and abs(abs(B[1] @ Z) - 1) < 1e-12
    assert np.allclose(cp, [0.5, 0.5], atol=1e-12)
    # point-edge
    B, cp = one([[0, 1, 0], [-1, 0, 0], [1, 0, 0]], [[1, 2]], [], "ev")
    assert abs(abs(B[0] @ X) - 1) < 1e-12 and abs(abs(B[1] @ Z) - 1) < 1e-12
    assert abs(cp[0] - 0.5) < 1e-12
    # point-point
    B, cp = one([[0, 0, 0], [0, 0, 1]], [], [], "vv")
    assert abs(abs(B[0] @ X) - 1) < 1e-12 and abs(abs(B[1] @ Y) - 1) < 1e-12


def check_reference_relative_velocity_kats(api):
    """test_relative_velocity.cpp:13-69: with the other primitive at rest the relative velocity IS the moving point's velocity
    (at the closest point), and a common translation gives none — observed through the one-collision friction potential
    E(u) = w N mu f0(|P^T u_rel|) with the reference's f0 (friction/smooth_friction_mollifier.cpp:7-14)"""
    eps_v, mu, dhat = 1e-2, 0.4, 2.0
    f0 = lambda y: y if abs(y) >= eps_v else y * y * (1 - y / (3 * eps_v)) / eps_v + eps_v / 3
    cases = [  # (V, E, F, kind, moving vertices: the first primitive)
        ([[0, 1, 0], [-1, 0, 1], [1, 0, 1], [0, 0, -1]], [[1, 2], [2, 3], [3, 1]], [[1, 2, 3]], "fv", [0]),
        ([[-1, 0, 0], [1, 0, 0], [0, 1, -1], [0, 1, 1]], [[0, 1], [2, 3]], [], "ee", [0, 1]),
        ([[0, 1, 0], [-1, 0, 0], [1, 0, 0]], [[1, 2]], [], "ev", [0]),
        ([[0, 0, 0], [0, 0, 1]], [], [], "vv", [0]),
    ]
    rng = np.random.default_rng(11)
    for V, E, F, key, moving in cases:
        V = np.asarray(V, float)
        mesh = api.CollisionMesh(V, np.asarray(E, np.int32).reshape(-1, 2), np.asarray(F, np.int32).reshape(-1, 3))
        c = api.NormalCollisions()
        c.build(mesh, V, dhat)
        t = api.TangentialCollisions()
        t.build(mesh, V, c, api.BarrierPotential(dhat, 1.0), np.full(len(V), mu), np.full(len(V), mu))
        r = getattr(t, key + "_collisions")
        assert t.size() == 1 and len(r.ids) == 1
        P, N, w = r.tangent_basis[0], r.normal_force_magnitude[0], r.weight[0]
        D = api.FrictionPotential(eps_v)
        for scale in (0.3 * eps_v, 5 * eps_v):  # sticking and sliding
            dp = rng.normal(0, 1, 3) * scale
            U = np.zeros_like(V)
            U[moving] = dp  # the whole first primitive moves with dp, the other one rests
            want = w * N * mu * f0(np.linalg.norm(P @ dp))
            assert D(t, mesh, U) == pytest.approx(want, rel=1e-11)
            U[:] = dp  # common translation: no relative velocity
            assert D(t, mesh, U) == pytest.approx(w * N * mu * eps_v / 3, rel=1e-11)
            assert np.abs(D.gradient(t, mesh, U)).max() <= 1e-12 * N


def test_reference_tangent_known_answers_oracle(oracle):
    check_reference_tangent_kats(oracle)
    check_reference_relative_velocity_kats(oracle)


@pytest.mark.gpu
def test_reference_tangent_known_answers(cuda):
    check_reference_tangent_kats(cuda)
    check_reference_relative_velocity_kats(cuda)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["stack", "drape", "soup"])
def test_friction_matches_the_oracle(cuda, oracle, scenes, name):
    V0, V1, E, F, P = _scene(scenes, name)
    ta, tb = check_lagged_quantities(cuda, scenes, name), check_lagged_quantities(oracle, scenes, name)
    assert ta.counts() == tb.counts()
    for key in ("vv", "ev", "ee", "fv"):
        a, b = getattr(ta, key + "_collisions"), getattr(tb, key + "_collisions")
        assert np.array_equal(a.ids, b.ids)
        for f in ("weight", "normal_force_magnitude", "mu_s", "mu_k", "closest_point", "tangent_basis"):
            x, y = getattr(a, f), getattr(b, f)
            assert np.allclose(x, y, rtol=1e-10, atol=1e-12 * max(1.0, np.abs(y).max() if y.size else 0.0)), (key, f)
    for eps_v in (1e-3, 1e-5):
        ra, rb = check_potential_fd(cuda, scenes, name, eps_v), check_potential_fd(oracle, scenes, name, eps_v)
        assert abs(ra["e"] - rb["e"]) <= 1e-10 * abs(rb["e"])
        assert np.linalg.norm(ra["g"] - rb["g"]) <= 1e-10 * np.linalg.norm(rb["g"])
        A, Bm = ra["H"], rb["H"]
        assert np.array_equal(A.indptr, Bm.indptr) and np.array_equal(A.indices, Bm.indices)
        assert np.linalg.norm(A.data - Bm.data) <= 1e-10 * np.linalg.norm(Bm.data)


@pytest.mark.gpu
def test_device_resident_friction_calls(cuda, scenes):
    """the _dev forms (device positions / velocities / coefficients in, results left on the device) give what the
    host-buffer forms give"""
    import ctypes as C

    import scipy.sparse as sp
    import torch

    V0, V1, E, F, P = _scene(scenes, "stack")
    nV, eps_v, lib = V0.shape[0], 1e-3, cuda.lib
    mesh, c, t = build_sets(cuda, V0, E, F, P["dhat"])
    rng = np.random.default_rng(5)
    U = rng.normal(0, eps_v, V0.shape)
    D = cuda.FrictionPotential(eps_v)
    want = dict(counts=t.counts(), e=D(t, mesh, U), g=D.gradient(t, mesh, U), H=D.hessian(t, mesh, U, cuda.PSDProjectionMethod.CLAMP))
    # the same through the device-resident entry points
    r2 = np.random.default_rng(1)  # build_sets' coefficients
    mu_s, mu_k = 0.5 * (1 + 0.2 * r2.random(nV)), 0.3 * (1 + 0.2 * r2.random(nV))
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    dV, dU = dev(np.asfortranarray(V0).T.copy()), dev(np.asfortranarray(U).T.copy())
    d_ms, d_mk = dev(mu_s), dev(mu_k)
    c._live()  # the normal set the lagged quantities come from must be the resident one
    import ipctk_b200

    bp = ipctk_b200._pkg._abi.BarrierParams(P["dhat"], 1e3, 0)
    counts, nnz = (C.c_int64 * 4)(), C.c_int64()
    ptr = lambda x: C.c_void_p(x.data_ptr())
    lib.check(lib.tangential_build_dev(mesh._ctx, ptr(dV), nV, C.byref(bp), ptr(d_ms), ptr(d_mk), counts))
    assert list(counts) == want["counts"]
    d_e, d_g = torch.zeros(1, dtype=torch.float64, device="cuda"), torch.zeros(3 * nV, dtype=torch.float64, device="cuda")
    lib.check(lib.friction_energy_dev(mesh._ctx, ptr(dU), nV, eps_v, ptr(d_e)))
    lib.check(lib.friction_gradient_dev(mesh._ctx, ptr(dU), nV, eps_v, ptr(d_g)))
    lib.check(lib.friction_hessian_dev(mesh._ctx, ptr(dU), nV, eps_v, 1, C.byref(nnz)))
    torch.cuda.synchronize()
    outer, inner, vals = np.zeros(3 * nV + 1, np.int32), np.zeros(nnz.value, np.int32), np.zeros(nnz.value)
    lib.check(lib.barrier_hessian_fetch(mesh._ctx, outer.ctypes.data_as(C.c_void_p), inner.ctypes.data_as(C.c_void_p), vals.ctypes.data_as(C.c_void_p)))
    H = sp.csc_matrix((vals, inner, outer), shape=(3 * nV, 3 * nV))
    assert abs(float(d_e.item()) - want["e"]) <= 1e-13 * abs(want["e"])
    assert np.linalg.norm(d_g.cpu().numpy() - want["g"]) <= 1e-13 * np.linalg.norm(want["g"])
    W = want["H"]
    assert np.array_equal(H.indptr, W.indptr) and np.array_equal(H.indices, W.indices)
    assert np.linalg.norm(H.data - W.data) <= 1e-13 * np.linalg.norm(W.data)
    del dV, dU, d_ms, d_mk, d_e, d_g
