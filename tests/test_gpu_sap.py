"""GPU sweep-and-prune broad phase (north_star subsystem 1; reference semantics broad_phase/sweep_and_prune.cpp:106-119):
the same predicate as the LBVH, hence bit-identical candidate sets — all six kinds, static and swept boxes, with and
without a collision filter, and through Candidates::build / NormalCollisions::build / the step size."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

KINDS = ("vertex_vertex", "edge_vertex", "edge_edge", "face_vertex", "edge_face", "face_face")


def _scenes(scenes):
    yield "stack", scenes.cloth_stack(3, 20)
    yield "drape", scenes.cloth_on_sphere(32, 16, drape=True)
    yield "soup", scenes.random_soup(120, seed=3)
    V0, V1, E, F, P = scenes.perturbed_sheets(3, 12)
    yield "sheets", (V0, V0 + 0.2 * (V1 - V0), E, F, P)


@pytest.mark.parametrize("swept", [False, True])
def test_sap_equals_lbvh_and_oracle_all_kinds(cuda, oracle, scenes, swept):
    for name, (V0, V1, E, F, P) in _scenes(scenes):
        res = {}
        for key in ("lbvh", "sap", "oracle"):
            api = oracle if key == "oracle" else cuda
            mesh = api.CollisionMesh(V0, E, F)
            if key == "oracle":
                oracle.set_broad_method(mesh, 1)
            else:
                mesh.set_broad_phase_method(key)
            bp = api.BroadPhase(mesh)
            bp.build(V0, V1 if swept else None, inflation_radius=0.5 * P["dhat"])
            res[key] = [getattr(bp, "detect_%s_candidates" % k)() for k in KINDS]
        for kind, a, b, c in zip(KINDS, res["sap"], res["lbvh"], res["oracle"]):
            assert np.array_equal(a, b) and np.array_equal(a, c), (name, kind, len(a), len(b), len(c))
            assert len(np.unique(a, axis=0)) == len(a)


def test_sap_with_a_collision_filter(cuda, oracle, scenes):
    V0, V1, E, F, P = scenes.cloth_stack(4, 10, gap=0.4)
    per_layer = V0.shape[0] // 4
    for api_key in ("sap", "oracle"):
        pass
    out = {}
    for key in ("lbvh", "sap"):
        mesh = cuda.CollisionMesh(V0, E, F)
        mesh.set_broad_phase_method(key)
        mesh.can_collide = cuda.make_vertex_patches_filter(np.arange(V0.shape[0]) // per_layer) & cuda.make_static_obstacle_filter(3 * per_layer)
        bp = cuda.BroadPhase(mesh)
        bp.build(V0, V1, inflation_radius=0.5 * P["dhat"])
        out[key] = [getattr(bp, "detect_%s_candidates" % k)() for k in KINDS]
    for a, b in zip(out["sap"], out["lbvh"]):
        assert np.array_equal(a, b)
    assert sum(len(a) for a in out["sap"]) > 0


def test_contact_step_with_sap(cuda, scenes):
    """Candidates::build, NormalCollisions::build, the potential and the step size do not care which broad phase ran"""
    V0, V1, E, F, P = scenes.cloth_on_sphere(32, 16, drape=True)
    res = {}
    for key in ("lbvh", "sap"):
        mesh = cuda.CollisionMesh(V0, E, F)
        mesh.set_broad_phase_method(key)
        cand = cuda.Candidates()
        cand.build(mesh, V0, V1, 0.0)
        swept = [np.asarray(cand.ee_candidates).copy(), np.asarray(cand.fv_candidates).copy()]
        c = cuda.NormalCollisions()
        c.build(mesh, V0, P["dhat"])
        B = cuda.BarrierPotential(P["dhat"], 1.0)
        res[key] = dict(swept=swept, ids=[getattr(c, k + "_collisions").ids.copy() for k in ("vv", "ev", "ee", "fv")], e=B(c, mesh, V0),
                        step=cuda.compute_collision_free_stepsize(mesh, V0, V1, narrow_phase_ccd=cuda.AdditiveCCD()))
    a, b = res["sap"], res["lbvh"]
    assert all(np.array_equal(x, y) for x, y in zip(a["swept"], b["swept"])) and all(np.array_equal(x, y) for x, y in zip(a["ids"], b["ids"]))
    assert a["e"] == b["e"] and a["step"] == pytest.approx(b["step"], rel=1e-12)
