"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on
identical seeded inputs.  Bars (BASELINE.json north_star): candidate and
collision sets bit-exact after canonical sort; Hessian sparsity identical;
energy / gradient / Hessian values within 1e-10 relative; CCD step never larger
than the oracle's by more than the tolerance."""
import numpy as np
import pytest

from ccd_tolerance import StepTolerance, check_step

pytestmark = pytest.mark.gpu

RTOL = 1e-10  # north_star: values within 1e-10 relative


def _scene(scenes, name):
    return {
        "c1": lambda: scenes.cloth_on_sphere(64, 32),
        "drape": lambda: scenes.cloth_on_sphere(48, 24, drape=True),
        "stack": lambda: scenes.cloth_stack(3, 30),
        "stack_tight": lambda: scenes.cloth_stack(4, 20, gap=0.2),
        "soup": lambda: scenes.random_soup(150, seed=7),
        "sheets": lambda: scenes.perturbed_sheets(3, 16),
        # Hessian columns with thousands of row blocks: block-per-column sort in shared memory / in global scratch
        "dense3": lambda: scenes.dense_sheet(16, 3.0),
        "dense6": lambda: scenes.dense_sheet(14, 6.0),
    }[name]()


SCENES = ["c1", "drape", "stack", "stack_tight", "soup", "sheets"]


def relerr(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    n = np.linalg.norm(b)
    return np.linalg.norm(a - b) / n if n > 0 else np.linalg.norm(a)


@pytest.mark.parametrize("name", SCENES)
def test_vertex_boxes_bit_exact(cuda, oracle, scenes, name):
    V0, V1, E, F, P = _scene(scenes, name)
    for swept in (False, True):
        res = []
        for api in (cuda, oracle):
            mesh = api.CollisionMesh(V0, E, F)
            bp = api.BroadPhase(mesh)
            bp.build(V0, V1 if swept else None, inflation_radius=0.5 * P["dhat"])
            res.append(bp.vertex_boxes())
        assert res[0].dtype == np.float32 and np.array_equal(res[0].view(np.uint32), res[1].view(np.uint32))


@pytest.mark.parametrize("name", SCENES)
@pytest.mark.parametrize("swept", [False, True])
@pytest.mark.parametrize("refit", ["range_query", "bottom_up"])
def test_broad_phase_all_kinds(cuda, oracle, scenes, name, swept, refit, monkeypatch):
    if refit == "bottom_up":  # node boxes by the arrival-counter refit instead of range queries over the sorted leaves
        if name not in ("stack", "soup"):
            pytest.skip("the alternative refit is covered on two scenes")
        monkeypatch.setenv("IPCB_REFIT_BOTTOM_UP", "1")
    V0, V1, E, F, P = _scene(scenes, name)
    if swept and name == "sheets":
        V1 = V0 + 0.2 * (V1 - V0)  # keep the 6-kind brute-force comparison small
    got, want = [], []
    for api, out in ((cuda, got), (oracle, want)):
        mesh = api.CollisionMesh(V0, E, F)
        if api is oracle:
            oracle.set_broad_method(mesh, 1)  # brute force over the float-box predicate
        bp = api.BroadPhase(mesh)
        bp.build(V0, V1 if swept else None, inflation_radius=0.5 * P["dhat"])
        for fn in ("vertex_vertex", "edge_vertex", "edge_edge", "face_vertex", "edge_face", "face_face"):
            out.append(getattr(bp, "detect_%s_candidates" % fn)())
    for kind, (a, b) in enumerate(zip(got, want)):
        assert a.shape == b.shape, "kind %d: %d vs %d candidates" % (kind, len(a), len(b))
        assert np.array_equal(a, b), "kind %d candidate sets differ" % kind
        assert len(np.unique(a, axis=0)) == len(a), "duplicate candidates"  # test_lbvh.cpp:142-149


@pytest.mark.parametrize("name", SCENES + ["dense3", "dense6"])
def test_collision_set_and_potential(cuda, oracle, scenes, name):
    V0, V1, E, F, P = _scene(scenes, name)
    dhat = P["dhat"]
    for area in (False, True):
        res = {}
        for key, api in (("cuda", cuda), ("oracle", oracle)):
            mesh = api.CollisionMesh(V0, E, F)
            c = api.NormalCollisions()
            c.set_use_area_weighting(area)
            c.build(mesh, V0, dhat)
            sets = [getattr(c, k + "_collisions") for k in ("vv", "ev", "ee", "fv")]
            B = api.BarrierPotential(dhat, 1.0, use_physical_barrier=area)
            X = V0 + 0.02 * dhat * np.sin(np.arange(V0.size).reshape(V0.shape))  # evaluate off the build point
            res[key] = dict(sets=sets, e=B(c, mesh, X), g=B.gradient(c, mesh, X),
                            h0=B.hessian(c, mesh, X), h1=B.hessian(c, mesh, X, api.PSDProjectionMethod.CLAMP),
                            h2=B.hessian(c, mesh, X, api.PSDProjectionMethod.ABS), dmin=c.compute_minimum_distance(mesh, X))
        a, b = res["cuda"], res["oracle"]
        assert sum(len(s.ids) for s in b["sets"]) > 0
        for sa, sb in zip(a["sets"], b["sets"]):
            assert np.array_equal(sa.ids, sb.ids)  # bit-exact collision set (canonical order)
            assert np.array_equal(sa.dtype, sb.dtype)
            assert np.array_equal(sa.eps_x, sb.eps_x)
            assert relerr(sa.weight, sb.weight) <= 1e-14
        assert abs(a["e"] - b["e"]) <= RTOL * abs(b["e"])
        assert relerr(a["g"], b["g"]) <= RTOL
        assert abs(a["dmin"] - b["dmin"]) <= 1e-14 * b["dmin"]
        for h in ("h0", "h1", "h2"):
            A, Bm = a[h], b[h]
            assert np.array_equal(A.indptr, Bm.indptr), h + ": outer index arrays differ"
            assert np.array_equal(A.indices, Bm.indices), h + ": sparsity pattern differs"
            assert relerr(A.data, Bm.data) <= RTOL, h
        # projected Hessians are PSD and symmetric
        H = a["h1"]
        assert abs(H - H.T).max() <= 1e-12 * abs(H).max()


@pytest.mark.parametrize("name", SCENES)
def test_step_size(cuda, oracle, scenes, name):
    V0, V1, E, F, P = _scene(scenes, name)
    for md in (0.0, 1e-4 * P["dhat"]):
        steps = {}
        for key, api in (("cuda", cuda), ("oracle", oracle)):
            mesh = api.CollisionMesh(V0, E, F)
            steps[key] = (api.compute_collision_free_stepsize(mesh, V0, V1, md),
                          api.compute_collision_free_stepsize(mesh, V0, V1, md, narrow_phase_ccd=api.AdditiveCCD()))
        (ti_g, ac_g), (ti_o, ac_o) = steps["cuda"], steps["oracle"]
        assert 0 <= ti_g <= 1 and 0 <= ac_g <= 1
        # Tight Inclusion: the bar is DERIVED (tests/ccd_tolerance.py) from the configured co-domain tolerance, the root
        # finder's floating-point filter and the closing speed measured on the scene: two valid searches answer with
        # lower time bounds of terminal boxes that can differ by the time the critical pair needs to close
        # sqrt(3) (delta + err) plus one box width.  It also checks, with the ORACLE's narrow phase, that a step by the
        # returned size (backed off by one box width) is collision free.
        check_step(oracle, V0, V1, E, F, ti_g, ti_o, md)
        # Additive CCD is the same arithmetic; the shared-bound pruning can only change which query stops first
        assert abs(ac_g - ac_o) <= 1e-9 * max(ac_o, 1e-12), (ac_g, ac_o)


@pytest.mark.parametrize("hooks", [
    {},  # hash path with fall-back to the sort path for columns with more than 128 unique row vertices
    {"IPCB_HESS_NO_HASH": "1"},  # warp sort path, block path for columns beyond 512 row blocks
    {"IPCB_HESS_NO_HASH": "1", "IPCB_HESS_WARP_CAP": "64"},  # most columns on the block-per-column kernel
    {"IPCB_HESS_NO_HASH": "1", "IPCB_HESS_WARP_CAP": "32", "IPCB_HESS_CTA_CAP": "128"},  # ... sorting in global scratch
    {"IPCB_HESS_NO_HASH": "1", "IPCB_HESS_WARP_CAP": "32", "IPCB_HESS_BIG_HASH": "2"},  # ... block-wide hash (giant columns)
    {"IPCB_HESS_RADIX_INCIDENCES": "1"},  # incidences grouped by the global radix sort instead of the counting placement
    {"IPCB_HESS_COLSORT_WARP_CAP": "16"},  # counting placement: most columns sorted by the block-per-column kernel
    {"IPCB_HESS_COLSORT_CTA_CAP": "8"},  # ... a column too large for it: the assembly is redone with the radix sort
    {"IPCB_NUMERIC_LANES": "3"},  # numeric pass: one lane per block column, ten blocks per round
    {"IPCB_NUMERIC_LANES": "3", "IPCB_NUM_BATCH": "4"},  # ... with runs longer than the prefetch depth
    {"IPCB_NUMERIC_LANES": "9", "IPCB_NUM_BATCH": "8"},  # one lane per block entry (the other variant)
    {"IPCB_NUM_REM": "0"},  # remainder of long runs shared by the three lane groups
    {"IPCB_NUM_COOP": "4"},  # ... four gathers in flight per lane
    {"IPCB_NUM_REM": "0", "IPCB_NUM_BATCH": "6"},  # ... behind a short first batch
    {"IPCB_NUM_REM": "4"},  # remainder of long runs four gathers at a time (default: one)
    {"IPCB_NUM_REM": "8"},  # ... eight at a time
    {"IPCB_HESS_NUMERIC_BIG": "40"},  # numeric pass: most columns on the block-per-column kernel
    {"IPCB_HESS_NUMERIC_BIG": "40", "IPCB_HESS_NUMERIC_UCAP": "7"},  # ... in several shared-memory segments per column
])
def test_hessian_column_paths(cuda, oracle, scenes, hooks, monkeypatch):
    """Hessian assembly: every per-column path (hash de-duplication in one warp, warp sort, block-per-column sort in
    shared memory, block sort in global scratch, block-wide hash de-duplication) must produce the oracle's matrix"""
    V0, V1, E, F, P = scenes.dense_sheet(14, 6.0)
    for k, v in hooks.items():
        monkeypatch.setenv(k, v)
    res = []
    for api in (cuda, oracle):
        mesh = api.CollisionMesh(V0, E, F)
        c = api.NormalCollisions()
        c.build(mesh, V0, P["dhat"])
        B = api.BarrierPotential(P["dhat"], 1.0)
        res.append([B.hessian(c, mesh, V0, m) for m in (api.PSDProjectionMethod.NONE, api.PSDProjectionMethod.CLAMP)])
    for A, Bm in zip(*res):
        assert np.array_equal(A.indptr, Bm.indptr) and np.array_equal(A.indices, Bm.indices)
        assert relerr(A.data, Bm.data) <= RTOL


def test_codim_known_answers(cuda):
    """collisions/test_normal_collisions.cpp:14-107 and :109-209 of the reference, through the CUDA path"""
    V = np.array([[0, 0, 0], [0, 0, 1], [0, 1, 0], [0, 1, 1], [1, 0, 0], [1, 0, 1], [1, 1, 0], [1, 1, 1]], float)
    V -= V.mean(0)
    mesh = cuda.CollisionMesh(V)
    assert mesh.num_codim_vertices() == 8
    c = cuda.NormalCollisions()
    c.build(mesh, V, 0.25, 0.8)
    assert c.counts() == [12, 0, 0, 0]
    B = cuda.BarrierPotential(0.25, 1.0)
    assert B(c, mesh, V) > 0
    f = -B.gradient(c, mesh, V).reshape(-1, 3)
    assert np.allclose(f / np.linalg.norm(f, axis=1, keepdims=True), V / np.linalg.norm(V, axis=1, keepdims=True))
    V1 = V.copy()
    V1[:, 1] *= 0.5
    cand = cuda.Candidates()
    cand.build(mesh, V, V1, 0.4)
    assert cand.size() == len(cand.vv_candidates) > 0
    assert not cand.is_step_collision_free(mesh, V, V1, 0.8)
    assert cand.compute_collision_free_stepsize(mesh, V, V1, 0.8) == pytest.approx((1 - (0.8 + 1e-4)) / 2 / 0.25, rel=1.2e-5)

    V = np.array([[0, 0, 0], [1, 0, 0], [0, 0, -1], [-1, 0, 0], [0, 0, 1], [0, 1, 0], [0, 2, 0], [0, 3, 0]], float)
    E = np.array([[0, 1], [0, 2], [0, 3], [0, 4]])
    mesh = cuda.CollisionMesh(V, E)
    assert (mesh.num_codim_vertices(), mesh.num_codim_edges()) == (3, 4)
    V1 = V.copy()
    V1[5:, 1] -= 4
    cand = cuda.Candidates()
    cand.build(mesh, V, V1, 1e-3)
    assert [len(cand.vv_candidates), len(cand.ev_candidates), len(cand.ee_candidates), len(cand.fv_candidates)] == [3, 12, 0, 0]
    assert cand.compute_collision_free_stepsize(mesh, V, V1, 2e-3) == pytest.approx((1 - (2e-3 + 1e-4)) / 4, rel=1.2e-5)
    c = cuda.NormalCollisions()
    c.build(mesh, V, 0.25, 0.8)
    assert c.counts() == [2, 4, 0, 0]


def test_narrow_phase_matches_oracle(cuda, oracle):
    rng = np.random.default_rng(5)
    for kind in (0, 1, 2, 3):
        n = 400 if kind >= 2 else 60
        a = rng.uniform(-1, 1, (n, 4, 3)) * (1.0 if kind >= 2 else 0.4)  # points / point-edge pairs need to start closer to meet
        b = a + rng.normal(0, 0.6, (n, 4, 3))
        # point-point / point-edge only meet with a real thickness; a coarser tolerance keeps the
        # sequential oracle fast there (the work grows with min_distance / tolerance)
        md = {0: 0.3, 1: 0.1, 2: 1e-3, 3: 1e-3}[kind]
        tol = 1e-6 if kind >= 2 else 1e-3
        for ccd in ("ti", "accd"):
            hg, tg = cuda.narrow_phase_ccd(kind, a, b, md, 1.0, cuda.AdditiveCCD() if ccd == "accd" else cuda.TightInclusionCCD(tol))
            ho, to = oracle.narrow_phase_ccd(kind, a, b, md, 1.0, oracle.AdditiveCCD() if ccd == "accd" else oracle.TightInclusionCCD(tol))
            assert np.array_equal(hg, ho), (kind, ccd, np.flatnonzero(hg != ho)[:5])
            assert hg.sum() > 0
            if ccd == "accd":
                assert np.allclose(tg[hg], to[ho], rtol=1e-9, atol=0)
            else:
                # no shared bound in this API: the GPU returns the earliest terminal box, the sequential
                # search the first one in (level, t) order — identical except when its width-only
                # stopping rule (condition 1) fires first, so: never later, and within the time tolerance
                assert np.all(tg[hg] <= to[ho] + 1e-12)
                assert np.all(tg[hg] >= to[ho] - 5e-3)
                assert np.mean(np.abs(tg[hg] - to[ho]) <= 1e-12) > 0.9


@pytest.mark.parametrize("hooks", [{"IPCB_TI_BUDGET": "1"}, {"IPCB_TI_BUDGET": "1", "IPCB_TI_WSTACK": "8"},
                                   {"IPCB_TI_SAMPLE": "16", "IPCB_TI_GROWTH": "4"}])
def test_ccd_later_stages(cuda, oracle, scenes, hooks, monkeypatch):
    """force every search into the warp-cooperative kernel (budget 1) and, with a tiny shared-memory stack, on into
    the global level-synchronous queue; with a tiny first sample the nested strided phases run on a small scene:
    the answers must not depend on which stage finishes a query"""
    rng = np.random.default_rng(9)
    a = rng.uniform(-1, 1, (300, 4, 3))
    b = a + rng.normal(0, 0.6, (300, 4, 3))
    V0, V1, E, F, P = scenes.cloth_stack(3, 30)
    ref_hits = {k: cuda.narrow_phase_ccd(k, a, b, 1e-3, 1.0) for k in (2, 3)}
    mesh = cuda.CollisionMesh(V0, E, F)
    ref_step = cuda.compute_collision_free_stepsize(mesh, V0, V1)
    for k, v in hooks.items():
        monkeypatch.setenv(k, v)
    for k in (2, 3):
        hit, toi = cuda.narrow_phase_ccd(k, a, b, 1e-3, 1.0)
        assert np.array_equal(hit, ref_hits[k][0]) and hit.sum() > 0
        assert np.array_equal(toi[hit], ref_hits[k][1][hit])  # order-independent minimum: bit-identical
        ho, to = oracle.narrow_phase_ccd(k, a, b, 1e-3, 1.0)
        assert np.array_equal(hit, ho)
    # the step size over a candidate SET is pruned with the shared earliest-TOI bound (candidates.cpp:267-286) and every
    # box is clipped to it, so — like the reference under TBB — the returned lower bound depends on which query tightens
    # the bound first: reproducible to the derived time tolerance (tests/ccd_tolerance.py), not bit for bit
    step = cuda.compute_collision_free_stepsize(mesh, V0, V1)
    tol = StepTolerance(oracle, V0, V1, E, F).tolerance(min(step, ref_step))[0]
    assert abs(step - ref_step) <= tol, (step, ref_step, tol)


def test_errors_are_reported(cuda):
    V = np.zeros((3, 3))
    with pytest.raises(RuntimeError, match="Unable to find edge!"):
        cuda.CollisionMesh(V, np.array([[0, 1]]), np.array([[0, 1, 2]]))  # collision_mesh.cpp:537
    mesh = cuda.CollisionMesh(np.eye(3))
    with pytest.raises(RuntimeError):
        cuda.BarrierPotential(1e-3)(cuda.NormalCollisions(), mesh, V)  # stale / unbuilt collision handle


def test_empty_inputs(cuda):
    V = np.array([[0.0, 0, 0], [1, 0, 0], [0, 1, 0]])
    mesh = cuda.CollisionMesh(V, np.array([[0, 1], [1, 2], [0, 2]]), np.array([[0, 1, 2]]))
    c = cuda.NormalCollisions()
    c.build(mesh, V, 1e-3)
    assert c.empty()
    B = cuda.BarrierPotential(1e-3)
    assert B(c, mesh, V) == 0.0
    assert not B.gradient(c, mesh, V).any()
    H = B.hessian(c, mesh, V)
    assert H.shape == (9, 9) and H.nnz == 0  # potential.cpp:107-109
    assert cuda.compute_collision_free_stepsize(mesh, V, V + 1.0) == 1.0  # candidates.cpp:263-265


@pytest.mark.parametrize("name", ["stack", "drape", "dense3"])
def test_collision_merge_and_sharded_potential(cuda, oracle, scenes, name):
    """NormalCollisionsBuilder::merge over several builders (collisions_clear / append / merge) and the sharded
    potential (collision ranges for energy / gradient, row blocks for the Hessian) against the oracle: merging the
    records of two overlapping halves restores the set with accumulated weights; the range / block results of three
    pretend ranks tile the single-context results."""
    import types

    V0, V1, E, F, P = _scene(scenes, name)
    dhat = P["dhat"]
    X = V0 + 0.02 * dhat * np.sin(np.arange(V0.size).reshape(V0.shape))
    world = 3
    out = {}
    for key, api in (("cuda", cuda), ("oracle", oracle)):
        mesh = api.CollisionMesh(V0, E, F)
        c = api.NormalCollisions()
        c.build(mesh, V0, dhat)
        full = [getattr(c, k + "_collisions") for k in ("vv", "ev", "ee", "fv")]
        B = api.BarrierPotential(dhat, 1.0)
        ref = dict(e=B(c, mesh, X), g=B.gradient(c, mesh, X), h=B.hessian(c, mesh, X, api.PSDProjectionMethod.CLAMP))

        def part(rec, sl, kind):  # VV / EV / EE records of the overlap appear in both builders; FV ones are never merged
            return types.SimpleNamespace(ids=rec.ids[sl], weight=rec.weight[sl], eps_x=rec.eps_x[sl], dtype=rec.dtype[sl])

        builders = []
        for b in range(2):
            kinds = []
            for kind, rec in enumerate(full):
                n = len(rec.ids)
                lo, hi = (0, (2 * n) // 3) if b == 0 else (n // 3, n)
                if kind == 3:
                    lo, hi = (0, n // 2) if b == 0 else (n // 2, n)
                kinds.append(part(rec, slice(lo, hi), kind))
            builders.append(kinds[::1])
        builders.reverse()  # order of the builders must not matter
        m = api.NormalCollisions()
        m.assign(mesh, builders, 0.0)
        merged = [getattr(m, k + "_collisions") for k in ("vv", "ev", "ee", "fv")]
        for kind, (a, b) in enumerate(zip(merged, full)):
            assert np.array_equal(a.ids, b.ids) and np.array_equal(a.dtype, b.dtype) and np.array_equal(a.eps_x, b.eps_x)
            n = len(b.ids)
            expect = b.weight.copy()
            if kind != 3:
                expect[n // 3:(2 * n) // 3] *= 2  # the overlap was contributed by both builders
            assert np.array_equal(a.weight, expect)
        # disjoint shards (IPCB_MERGE_DISJOINT_SHARDS): EE / FV records are concatenated, VV / EV united; the canonical
        # order comes back when the records are fetched
        halves = [[part(rec, slice(b, None, 2), kind) for kind, rec in enumerate(full)] for b in (1, 0)]
        m.assign(mesh, halves, 0.0, disjoint_shards=True)
        e_concat = B(m, mesh, X)
        assert abs(e_concat - ref["e"]) <= 1e-12 * abs(ref["e"])
        for a, b in zip([getattr(m, k + "_collisions") for k in ("vv", "ev", "ee", "fv")], full):
            assert np.array_equal(a.ids, b.ids) and np.array_equal(a.dtype, b.dtype) and np.array_equal(a.eps_x, b.eps_x)
            assert np.array_equal(a.weight, b.weight)
        # back to one builder, then the sharded potential
        m.assign(mesh, [full], 0.0)
        bounds = mesh.balanced_row_blocks(world)
        e, g, tiles = 0.0, 0.0, []
        for r in range(world):
            mesh.set_collision_range(r, world)
            mesh.set_row_block(int(bounds[r]), int(bounds[r + 1]))
            e += B(m, mesh, X)
            g = g + B.gradient(m, mesh, X)
            tiles.append(B.hessian(m, mesh, X, api.PSDProjectionMethod.CLAMP))
        mesh.set_collision_range(0, 1)
        mesh.set_row_block()
        assert abs(e - ref["e"]) <= 1e-12 * abs(ref["e"]) and relerr(g, ref["g"]) <= 1e-12
        H = ref["h"].tocsc()
        for r, T in enumerate(tiles):
            lo, hi = 3 * int(bounds[r]), 3 * int(bounds[r + 1])
            T = T.tocsc()
            assert T[:, :lo].nnz == 0 and T[:, hi:].nnz == 0
            A, Bm = T[:, lo:hi], H[:, lo:hi]
            assert np.array_equal(A.indptr, Bm.indptr) and np.array_equal(A.indices, Bm.indices)
            assert relerr(A.data, Bm.data) <= 1e-13
        assert sum(T.nnz for T in tiles) == H.nnz
        out[key] = dict(bounds=bounds, merged=merged, tiles=tiles)
    assert np.array_equal(out["cuda"]["bounds"], out["oracle"]["bounds"])
    assert out["cuda"]["bounds"][0] == 0 and out["cuda"]["bounds"][-1] == V0.shape[0]
    for a, b in zip(out["cuda"]["tiles"], out["oracle"]["tiles"]):
        assert np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices) and relerr(a.data, b.data) <= RTOL


@pytest.mark.parametrize("name", ["stack", "drape", "sheets"])
def test_line_search_rebuild_from_resident_candidates(cuda, oracle, scenes, name):
    """SURVEY §8f rank 2 — the caller loop of the reference's solver example (python/examples/solver.py:95-116):
    candidates built ONCE over the swept step with inflation dhat, the collision-free step size from them, then
    NormalCollisions::build(candidates, mesh, X, dhat) + the barrier energy at several line-search points, all from the
    resident candidate set (no broad phase in the loop).  Sets bit-exact, energies / minimum distances to 1e-10."""
    V0, V1, E, F, P = _scene(scenes, name)
    dhat = P["dhat"]
    state = {}
    for key, api in (("cuda", cuda), ("oracle", oracle)):
        mesh = api.CollisionMesh(V0, E, F)
        cand = api.Candidates()
        cand.build(mesh, V0, V1, dhat)
        alpha = cand.compute_collision_free_stepsize(mesh, V0, V1)
        state[key] = dict(api=api, mesh=mesh, cand=cand, alpha=alpha, n=[len(cand.vv_candidates), len(cand.ev_candidates),
                                                                        len(cand.ee_candidates), len(cand.fv_candidates)],
                          ee=cand.ee_candidates.copy(), fv=cand.fv_candidates.copy())
    a, b = state["cuda"], state["oracle"]
    assert a["n"] == b["n"] and np.array_equal(a["ee"], b["ee"]) and np.array_equal(a["fv"], b["fv"])
    # the candidates were built with inflation dhat: the derived tolerance is evaluated on the same (larger) candidate set
    tol = StepTolerance(oracle, V0, V1, E, F, min_distance=2 * dhat).tolerance(min(a["alpha"], b["alpha"]))[0]
    assert abs(a["alpha"] - b["alpha"]) <= tol, (a["alpha"], b["alpha"], tol)
    alpha = min(a["alpha"], b["alpha"])
    B = {k: s["api"].BarrierPotential(dhat, 1.0) for k, s in state.items()}
    energies = []
    for it in range(4):  # backtracking: alpha, alpha / 2, ...
        X = V0 + alpha * 0.5 ** it * (V1 - V0)
        res = {}
        for key, s in state.items():
            c = s["api"].NormalCollisions()
            c.build(s["cand"], s["mesh"], X, dhat)
            sets = [getattr(c, k + "_collisions") for k in ("vv", "ev", "ee", "fv")]
            res[key] = dict(sets=sets, e=B[key](c, s["mesh"], X), d=c.compute_minimum_distance(s["mesh"], X),
                            g=B[key].gradient(c, s["mesh"], X))
        for sa, sb in zip(res["cuda"]["sets"], res["oracle"]["sets"]):
            assert np.array_equal(sa.ids, sb.ids) and np.array_equal(sa.dtype, sb.dtype) and relerr(sa.weight, sb.weight) <= 1e-14
        assert abs(res["cuda"]["e"] - res["oracle"]["e"]) <= RTOL * abs(res["oracle"]["e"])
        assert relerr(res["cuda"]["g"], res["oracle"]["g"]) <= RTOL
        if np.isfinite(res["oracle"]["d"]):
            assert abs(res["cuda"]["d"] - res["oracle"]["d"]) <= 1e-12 * res["oracle"]["d"]
        else:
            assert not np.isfinite(res["cuda"]["d"])
        energies.append(res["cuda"]["e"])
    assert any(e > 0 for e in energies)


def test_unsupported_set_types_fail_loudly(cuda, scenes):
    """OGC is outside the path; IMPROVED_MAX_APPROX on a sharded context needs the exchange of the sub-element keys
    (collisions_corrections_*_dev, tests/test_zz_improved_max_approx_gpu.py), so the host-buffer build must refuse it
    instead of silently building something else"""
    V0, V1, E, F, P = _scene(scenes, "stack")
    mesh = cuda.CollisionMesh(V0, E, F)
    c = cuda.NormalCollisions()
    c.set_collision_set_type(cuda.NormalCollisions.CollisionSetType.OGC)
    with pytest.raises(NotImplementedError):
        c.build(mesh, V0, P["dhat"])
    cuda.lib.check(cuda.lib.ctx_set_shard(mesh._ctx, 0, 2))
    c.set_collision_set_type(cuda.NormalCollisions.CollisionSetType.IMPROVED_MAX_APPROX)
    with pytest.raises(RuntimeError, match="IMPROVED_MAX_APPROX"):
        c.build(mesh, V0, P["dhat"])
    cuda.lib.check(cuda.lib.ctx_set_shard(mesh._ctx, 0, 1))
    c.set_collision_set_type(cuda.NormalCollisions.CollisionSetType.IPC)
    c.build(mesh, V0, P["dhat"])
    assert sum(c.counts()) > 0


EE_SWAP = np.array([0, 2, 1, 3, 6, 7, 4, 5, 8], np.uint8)  # distance type of (eb, ea) given the type of (ea, eb)


def test_edge_edge_orientation_is_never_lost(cuda, oracle, scenes):
    """(1) user-set edge-edge candidates given as (max, min) are classified like (min, max) ones; (2) collision
    records appended with edges stored as (max, min) keep that orientation AND their distance type through append / merge /
    lazy sort, and the potential evaluates them like the mirrored record."""
    V0, V1, E, F, P = _scene(scenes, "stack_tight")
    dhat = P["dhat"]
    X = V0 + 0.02 * dhat * np.sin(np.arange(V0.size).reshape(V0.shape))
    out = {}
    for key, api in (("cuda", cuda), ("oracle", oracle)):
        mesh = api.CollisionMesh(V0, E, F)
        cand = api.Candidates()
        cand.build(mesh, V0, 0.5 * dhat)
        ee, fv = np.asarray(cand.ee_candidates).copy(), np.asarray(cand.fv_candidates).copy()
        c = api.NormalCollisions()
        c.build(cand, mesh, V0, dhat)
        ref = [getattr(c, k + "_collisions") for k in ("vv", "ev", "ee", "fv")]
        B = api.BarrierPotential(dhat, 1.0)
        e_ref, g_ref = B(c, mesh, X), B.gradient(c, mesh, X)
        # (1) every other candidate reversed
        mixed = ee.copy()
        mixed[::2] = mixed[::2, ::-1]
        c2 = api.Candidates()
        c2.set(mesh, ee=mixed, fv=fv)
        c = api.NormalCollisions()
        c.build(c2, mesh, V0, dhat)
        for k, r in zip(("vv", "ev", "ee", "fv"), ref):
            s = getattr(c, k + "_collisions")
            assert np.array_equal(s.ids, r.ids) and np.array_equal(s.dtype, r.dtype) and np.array_equal(s.weight, r.weight), (key, k)
        # (2) records with every third edge pair stored as (max, min) and the distance type re-expressed for that order
        import types

        flipped = types.SimpleNamespace(ids=ref[2].ids.copy(), weight=ref[2].weight.copy(), eps_x=ref[2].eps_x.copy(), dtype=ref[2].dtype.copy())
        flipped.ids[::3] = flipped.ids[::3, ::-1]
        flipped.dtype[::3] = EE_SWAP[flipped.dtype[::3]]
        c = api.NormalCollisions()
        c.assign(mesh, [(ref[0], ref[1], flipped, ref[3])])
        got = c.ee_collisions
        order = np.lexsort((np.maximum(flipped.ids[:, 0], flipped.ids[:, 1]), np.minimum(flipped.ids[:, 0], flipped.ids[:, 1])))
        assert np.array_equal(got.ids, flipped.ids[order]) and np.array_equal(got.dtype, flipped.dtype[order]), key
        e, g = B(c, mesh, X), B.gradient(c, mesh, X)
        assert abs(e - e_ref) <= 1e-12 * abs(e_ref) and relerr(g, g_ref) <= 1e-12, key
        out[key] = (e, got.ids.copy(), got.dtype.copy())
    assert np.array_equal(out["cuda"][1], out["oracle"][1]) and np.array_equal(out["cuda"][2], out["oracle"][2])
    assert abs(out["cuda"][0] - out["oracle"][0]) <= RTOL * abs(out["oracle"][0])


@pytest.mark.parametrize("limit", [4000, 1500])
def test_streaming_step_size_when_the_candidate_list_is_capped(cuda, oracle, scenes, limit, monkeypatch):
    """SURVEY §7 hard part 7 / BASELINE config 4: more swept candidates than one list may hold (IPCB_MAX_PAIRS, here tiny)
    -> compute_collision_free_stepsize traverses the query leaves in chunks and pushes every chunk through the narrow
    phase; static candidate / collision builds beyond the limit fail loudly"""
    V0, V1, E, F, P = _scene(scenes, "sheets")
    ref = {}
    for ccd in ("ti", "accd"):
        mesh = cuda.CollisionMesh(V0, E, F)
        ref[ccd] = cuda.compute_collision_free_stepsize(mesh, V0, V1, narrow_phase_ccd=cuda.AdditiveCCD() if ccd == "accd" else None)
    mesh = cuda.CollisionMesh(V0, E, F)
    cand = cuda.Candidates()
    cand.build(mesh, V0, V1, 0.0)
    assert len(cand.ee_candidates) > 4 * 4000  # the limits below really force chunks
    monkeypatch.setenv("IPCB_MAX_PAIRS", str(limit))
    mesh = cuda.CollisionMesh(V0, E, F)
    ac = cuda.compute_collision_free_stepsize(mesh, V0, V1, narrow_phase_ccd=cuda.AdditiveCCD())
    assert abs(ac - ref["accd"]) <= 1e-12 * ref["accd"]  # same arithmetic per candidate, minimum over the same set
    ti = cuda.compute_collision_free_stepsize(mesh, V0, V1)
    tol = StepTolerance(oracle, V0, V1, E, F).tolerance(min(ti, ref["ti"]))[0]
    assert abs(ti - ref["ti"]) <= tol
    with pytest.raises(RuntimeError, match="IPCB_MAX_PAIRS"):
        cuda.Candidates().build(mesh, V0, V1, 0.0)
    monkeypatch.delenv("IPCB_MAX_PAIRS")
    assert cuda.compute_collision_free_stepsize(mesh, V0, V1, narrow_phase_ccd=cuda.AdditiveCCD()) == pytest.approx(ref["accd"], rel=1e-12)
