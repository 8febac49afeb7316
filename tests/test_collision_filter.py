"""CollisionMesh::can_collide / BroadPhase::can_vertices_collide (reference: collision_filter.hpp:30-143,
broad_phase.cpp:127-202, candidates.cpp:61,107,147) through the descriptor of the C ABI
(ipcb_mesh_set_collision_filter): vertex patches, static obstacles and their intersection.

CPU suite: the oracle's filtered candidate sets equal the unfiltered sets restricted by the reference's predicate
("no shared vertex and SOME pair of vertices of the two primitives can collide"), evaluated here in numpy.
GPU suite: the CUDA sets equal the oracle's, bit for bit, for all six candidate kinds, static and swept, and through
NormalCollisions::build / compute_collision_free_stepsize — including the codimensional passes, where the reference
(and therefore both libraries) evaluates the filter on the ids of the re-indexed vertex subsets.
"""
import numpy as np
import pytest

KINDS = ("vertex_vertex", "edge_vertex", "edge_edge", "face_vertex", "edge_face", "face_face")


def _filters(api, scenes):
    """(name, scene, filter) triples"""
    out = []
    V0, V1, E, F, P = scenes.cloth_on_sphere(20, 12, drape=True)
    n_cloth = P["n_cloth_vertices"]
    out.append(("static_obstacle", (V0, V1, E, F, P), api.make_static_obstacle_filter(n_cloth)))
    out.append(("components", (V0, V1, E, F, P), api.make_connected_components_filter(F, V0.shape[0])))
    V0, V1, E, F, P = scenes.cloth_stack(4, 10, gap=0.4)
    per_layer = V0.shape[0] // 4
    patches = np.arange(V0.shape[0]) // per_layer  # one patch per sheet: no self-contact inside a sheet
    out.append(("patches", (V0, V1, E, F, P), api.make_vertex_patches_filter(patches)))
    # sheets 0 and 1 share a patch, sheets 2 and 3 are static obstacles
    both = api.make_vertex_patches_filter(np.minimum(patches, 1) * 0 + (patches >= 1)) & api.make_static_obstacle_filter(2 * per_layer)
    out.append(("patches&static", (V0, V1, E, F, P), both))
    return out


def _prim_vertices(kind, E, F):
    a = {"vertex": None, "edge": E, "face": F}
    left, right = kind.split("_")
    return a[left], a[right]


def _expected(kind, pairs, E, F, f):
    """the reference predicate on a list of candidate pairs"""
    A, B = _prim_vertices(kind, E, F)
    keep = []
    for x, y in pairs:
        va = [x] if A is None else list(A[x])
        vb = [y] if B is None else list(B[y])
        shared = any(i == j for i in va for j in vb)
        keep.append((not shared) and any(f(i, j) for i in va for j in vb))
    return pairs[np.asarray(keep, bool)] if len(pairs) else pairs


def _all_kinds(api, mesh, V0, V1, r):
    bp = api.BroadPhase(mesh)
    bp.build(V0, V1, inflation_radius=r)
    return [getattr(bp, "detect_%s_candidates" % k)() for k in KINDS]


def test_filter_descriptor_algebra(oracle):
    f = oracle.make_vertex_patches_filter([0, 0, 1, 1])
    g = oracle.make_static_obstacle_filter(2)
    assert not f(0, 1) and f(1, 2) and g(0, 3) and not g(2, 3)
    h = f & g
    assert h(1, 2) and not h(0, 1) and not h(2, 3)
    assert (g & oracle.make_static_obstacle_filter(1)).n_dynamic == 1
    with pytest.raises(NotImplementedError):
        f & f
    cc = oracle.make_connected_components_filter(np.array([[0, 1, 2], [3, 4, 5]]))
    assert not cc(0, 2) and cc(0, 3)


def test_oracle_filter_matches_the_reference_predicate(oracle, scenes):
    for name, (V0, V1, E, F, P), f in _filters(oracle, scenes):
        mesh = oracle.CollisionMesh(V0, E, F)
        oracle.set_broad_method(mesh, 1)
        r = 0.5 * P["dhat"]
        for swept in (False, True):
            free = _all_kinds(oracle, mesh, V0, V1 if swept else None, r)
            mesh.can_collide = f
            got = _all_kinds(oracle, mesh, V0, V1 if swept else None, r)
            mesh.can_collide = oracle.CollisionFilter()
            removed = 0
            for kind, a, b in zip(KINDS, got, free):
                want = _expected(kind, b, E, F, f)
                assert np.array_equal(a, want), (name, kind, swept)
                removed += len(b) - len(a)
            assert removed > 0, name  # the filter did something


def _codim_scene():
    """codimensional vertices and edges around a small closed surface (a tetrahedron)"""
    rng = np.random.default_rng(3)
    tet = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], float)
    F = np.array([[0, 2, 1], [0, 1, 3], [0, 3, 2], [1, 2, 3]], np.int32)
    pts = rng.uniform(-0.2, 1.2, (14, 3))  # codim vertices
    seg = rng.uniform(-0.2, 1.2, (12, 3))  # 6 codim edges
    V = np.concatenate([tet, pts, seg])
    E_tet = np.array([[0, 1], [0, 2], [0, 3], [1, 2], [1, 3], [2, 3]], np.int32)
    E_seg = 18 + np.arange(12, dtype=np.int32).reshape(-1, 2)
    E = np.concatenate([E_tet, E_seg])
    V1 = V + rng.normal(0, 0.2, V.shape)
    return V, V1, E, F


def _candidate_lists(api, mesh, V0, V1, r):
    c = api.Candidates()
    if V1 is None:
        c.build(mesh, V0, r)
    else:
        c.build(mesh, V0, V1, r)
    return [np.asarray(x).copy() for x in (c.vv_candidates, c.ev_candidates, c.ee_candidates, c.fv_candidates)]


def test_codim_passes_see_local_ids_oracle(oracle):
    """candidates.cpp:61-77,83-108: the codimensional passes run on re-indexed vertex subsets with the SAME filter
    object, i.e. the filter is evaluated on positions in those subsets"""
    V, V1, E, F = _codim_scene()
    mesh = oracle.CollisionMesh(V, E, F)
    assert (mesh.num_codim_vertices(), mesh.num_codim_edges()) == (14, 6)
    free = _candidate_lists(oracle, mesh, V, None, 0.3)
    assert len(free[0]) > 5 and len(free[1]) > 5
    patches = np.arange(V.shape[0]) % 3
    mesh.can_collide = oracle.make_vertex_patches_filter(patches)
    got = _candidate_lists(oracle, mesh, V, None, 0.3)
    cv = np.arange(4, 18)  # codim vertices in ascending order: local id = position
    local = {int(v): i for i, v in enumerate(cv)}
    want_vv = np.array([p for p in free[0] if patches[local[p[0]]] != patches[local[p[1]]]], np.int32).reshape(-1, 2)
    assert np.array_equal(got[0], want_vv) and 0 < len(want_vv) < len(free[0])
    ref = {int(v): 14 + i for i, v in enumerate(np.arange(18, 30))}  # referenced vertices of the codim edges, ascending
    keep = [patches[local[v]] != patches[ref[E[e, 0]]] or patches[local[v]] != patches[ref[E[e, 1]]] for e, v in free[1]]
    assert np.array_equal(got[1], free[1][np.asarray(keep, bool)])
    # a filter that blocks one endpoint's patch AND the other's: labels 0 for every codim-edge vertex and most codim vertices
    lab = np.ones(V.shape[0], np.int32)
    lab[:7] = 0  # local ids 0..6 (codim vertices 4..10) ...
    lab[14:26] = 0  # ... share the label of the local ids of all codim-edge vertices
    mesh.can_collide = oracle.make_vertex_patches_filter(lab)
    got = _candidate_lists(oracle, mesh, V, None, 0.3)
    want = free[1][np.asarray([lab[local[v]] != 0 for e, v in free[1]], bool)]
    assert np.array_equal(got[1], want) and 0 < len(want) < len(free[1])


@pytest.mark.gpu
def test_filtered_candidates_match_the_oracle(cuda, oracle, scenes):
    for name, (V0, V1, E, F, P), _ in _filters(cuda, scenes):
        f_by = {id(cuda): dict((n, f) for n, _, f in _filters(cuda, scenes))[name], id(oracle): dict((n, f) for n, _, f in _filters(oracle, scenes))[name]}
        r = 0.5 * P["dhat"]
        for swept in (False, True):
            res = []
            for api in (cuda, oracle):
                mesh = api.CollisionMesh(V0, E, F)
                if api is oracle:
                    oracle.set_broad_method(mesh, 1)
                mesh.can_collide = f_by[id(api)]
                res.append(_all_kinds(api, mesh, V0, V1 if swept else None, r))
            for kind, a, b in zip(KINDS, *res):
                assert np.array_equal(a, b), (name, kind, swept)


@pytest.mark.gpu
def test_filtered_contact_step_matches_the_oracle(cuda, oracle, scenes):
    """the filter reaches NormalCollisions::build(mesh, V, dhat) and compute_collision_free_stepsize"""
    for name in ("static_obstacle", "patches&static"):
        out = {}
        for key, api in (("cuda", cuda), ("oracle", oracle)):
            (V0, V1, E, F, P), f = [(s, f) for n, s, f in _filters(api, scenes) if n == name][0]
            mesh = api.CollisionMesh(V0, E, F)
            c = api.NormalCollisions()
            c.build(mesh, V0, P["dhat"])
            free = c.counts()
            mesh.can_collide = f
            c = api.NormalCollisions()
            c.build(mesh, V0, P["dhat"])
            sets = [getattr(c, k + "_collisions") for k in ("vv", "ev", "ee", "fv")]
            B = api.BarrierPotential(P["dhat"], 1.0)
            out[key] = dict(free=free, counts=c.counts(), ids=[s.ids.copy() for s in sets], e=B(c, mesh, V0),
                            step=api.compute_collision_free_stepsize(mesh, V0, V1, narrow_phase_ccd=api.AdditiveCCD()))
        a, b = out["cuda"], out["oracle"]
        assert a["counts"] == b["counts"] and a["free"] == b["free"], name
        # the static-obstacle filter only removes sphere-sphere candidates (no collisions among them in this scene);
        # the patch filter removes the contacts between the two sheets that share a patch
        assert sum(a["counts"]) < sum(a["free"]) if name == "patches&static" else sum(a["counts"]) <= sum(a["free"]), name
        for x, y in zip(a["ids"], b["ids"]):
            assert np.array_equal(x, y)
        assert abs(a["e"] - b["e"]) <= 1e-10 * abs(b["e"])
        assert abs(a["step"] - b["step"]) <= 1e-9 * b["step"]


@pytest.mark.gpu
def test_codim_passes_with_a_filter_match_the_oracle(cuda, oracle):
    V, V1, E, F = _codim_scene()
    patches = np.arange(V.shape[0]) % 3
    for swept in (False, True):
        res = []
        for api in (cuda, oracle):
            mesh = api.CollisionMesh(V, E, F)
            mesh.can_collide = api.make_vertex_patches_filter(patches) & api.make_static_obstacle_filter(25)
            res.append(_candidate_lists(api, mesh, V, V1 if swept else None, 0.3))
        for kind, (a, b) in enumerate(zip(*res)):
            assert np.array_equal(a, b), (kind, swept)
        assert len(res[1][0]) > 0 and len(res[1][1]) > 0
