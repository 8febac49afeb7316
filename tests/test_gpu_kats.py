"""GPU side of the reference's known-answer tests and of the committed golden fixtures, through the C ABI.

* the check_* functions of tests/test_reference_kats.py (reference test cases with analytic expectations) run on the
  CUDA path;
* the distance-type vectors of tests/golden/reference_kats.npz are pushed through tiny meshes: the collision that
  NormalCollisions::build produces for a single candidate IS its distance type (normal_collisions_builder.cpp:140-336);
* tests/golden/scene_*.npz (oracle outputs frozen by tests/golden/make_golden.py) are compared bit-exactly (sets,
  sparsity) / to 1e-10 (values) without any oracle on the GPU box.
"""
import os

import numpy as np
import pytest

import kats
import test_reference_kats as ref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1e-10


def relerr(a, b):
    n = np.linalg.norm(b)
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / n if n > 0 else np.linalg.norm(a)


@pytest.mark.gpu
def test_broad_phase_kats_gpu(cuda):
    ref.check_broad_phase_kats(cuda)


@pytest.mark.gpu
def test_codim_kats_gpu(cuda):
    ref.check_codim_kats(cuda)
    ref.check_faces_to_edges_kats(cuda)


@pytest.mark.gpu
def test_barrier_potential_scenes_gpu(cuda):
    ref.check_barrier_potential_scenes(cuda)


@pytest.mark.gpu
def test_readme_quick_start_gpu(cuda):
    ref.check_readme_quick_start(cuda)


@pytest.mark.gpu
def test_ccd_kats_gpu(cuda):
    ref.check_ccd_kats(cuda)


def pt_type_via_pipeline(api, x):
    """point-triangle distance type of one stencil, read off the collision built from its single candidate"""
    V = x.reshape(4, 3)
    F = np.array([[1, 2, 3]], np.int32)
    E = np.array([[1, 2], [2, 3], [3, 1]], np.int32)  # faces_to_edges order == triangle edge order
    mesh = api.CollisionMesh(V, E, F)
    c = api.NormalCollisions()
    c.build(mesh, V, 1e6)  # dhat large: every candidate is active
    n = c.counts()
    assert sum(n) == 1
    if n[3]:
        return kats.P_T
    if n[1]:
        e = int(c.ev_collisions.ids[0][0])
        return kats.PT_E0 + int(mesh.faces_to_edges()[0].tolist().index(e))
    a, b = c.vv_collisions.ids[0]
    return kats.P_T0 + (max(a, b) - 1)


def ee_type_via_pipeline(api, x):
    V = x.reshape(4, 3)
    E = np.array([[0, 1], [2, 3]], np.int32)
    mesh = api.CollisionMesh(V, E)
    c = api.NormalCollisions()
    c.build(mesh, V, 1e6)
    n = c.counts()
    assert sum(n) == 1
    if n[2]:
        return int(c.ee_collisions.dtype[0])  # the collision keeps the actual type (edge_edge.hpp:96)
    if n[1]:
        e, v = c.ev_collisions.ids[0]
        return (kats.EA_EB0 + (v - 2)) if e == 0 else (kats.EA0_EB + v)
    a, b = sorted(c.vv_collisions.ids[0])
    return 2 * a + (b - 2)


def _sample(n, k, seed):
    return np.random.default_rng(seed).choice(n, size=min(k, n), replace=False)


def check_types_via_pipeline(api, n_pt=250, n_ee=250):
    g = np.load(os.path.join(GOLDEN, "reference_kats.npz"))
    for i in _sample(len(g["pt_x"]), n_pt, 1):
        assert pt_type_via_pipeline(api, g["pt_x"][i]) == g["pt_type"][i], i
    assert pt_type_via_pipeline(api, g["pt_x"][-1]) == kats.PT_E0  # the GH issue vector
    for i in _sample(len(g["ee_x"]), n_ee, 2):
        assert g["ee_ok"][i][ee_type_via_pipeline(api, g["ee_x"][i])], i


def test_types_via_pipeline_oracle(oracle):
    check_types_via_pipeline(oracle)


@pytest.mark.gpu
def test_types_via_pipeline_gpu(cuda):
    check_types_via_pipeline(cuda)


SCENES = ["c1_small", "drape", "stack", "soup"]


def check_golden_scene(api, name, oracle_for_tolerance=None):
    g = np.load(os.path.join(GOLDEN, "scene_%s.npz" % name))
    V0, V1, E, F, dhat, X = g["V0"], g["V1"], g["E"], g["F"], float(g["dhat"]), g["X"]
    mesh = api.CollisionMesh(V0, E, F)
    cand = api.Candidates()
    cand.build(mesh, V0, 0.5 * dhat)
    assert np.array_equal(np.asarray(cand.ee_candidates).reshape(-1, 2), g["ee_cand"].reshape(-1, 2))
    assert np.array_equal(np.asarray(cand.fv_candidates).reshape(-1, 2), g["fv_cand"].reshape(-1, 2))
    c = api.NormalCollisions()
    c.build(mesh, V0, dhat)
    for k in ("vv", "ev", "ee", "fv"):
        s = getattr(c, k + "_collisions")
        assert np.array_equal(np.asarray(s.ids).reshape(-1, 2), g[k + "_ids"].reshape(-1, 2)), k
        assert np.allclose(s.weight, g[k + "_w"], rtol=1e-14, atol=0)
    assert np.array_equal(c.ee_collisions.dtype, g["ee_dtype"]) and np.array_equal(c.ee_collisions.eps_x, g["ee_eps"])
    B = api.BarrierPotential(dhat, 1.0)
    assert abs(B(c, mesh, X) - float(g["energy"])) <= RTOL * abs(float(g["energy"]))
    assert relerr(B.gradient(c, mesh, X), g["gradient"]) <= RTOL
    for mode in (0, 1, 2):
        H = B.hessian(c, mesh, X, mode)
        assert np.array_equal(H.indptr, g["h%d_indptr" % mode]) and np.array_equal(H.indices, g["h%d_indices" % mode]), mode
        assert relerr(H.data, g["h%d_data" % mode]) <= RTOL, mode
    ti = api.compute_collision_free_stepsize(mesh, V0, V1)
    ac = api.compute_collision_free_stepsize(mesh, V0, V1, narrow_phase_ccd=api.AdditiveCCD())
    if oracle_for_tolerance is None:  # the oracle reproduces its own fixture
        assert ti == float(g["step_ti"])
    else:  # derived tolerance + the oracle's own collision-free check (tests/ccd_tolerance.py)
        from ccd_tolerance import check_step

        check_step(oracle_for_tolerance, V0, V1, E, F, ti, float(g["step_ti"]))
    assert abs(ac - float(g["step_accd"])) <= 1e-9 * float(g["step_accd"])


@pytest.mark.parametrize("name", SCENES)
def test_golden_scene_oracle(oracle, name):
    """the committed fixtures are reproducible from the oracle (they were generated by it)"""
    check_golden_scene(oracle, name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", SCENES)
def test_golden_scene_gpu(cuda, oracle, name):
    check_golden_scene(cuda, name, oracle)


@pytest.mark.gpu
def test_independent_collision_sets_gpu(cuda, scenes):
    ref.check_independent_collision_sets(cuda, scenes)
