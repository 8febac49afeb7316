"""ipc::has_intersections (reference ipc.cpp:105-166, geometry/intersection.cpp:115-145): edge-face
candidates + exact orientation + LU solve.  CPU suite: the oracle (exact integer arithmetic for the orientation) on
constructed cases; GPU suite: the CUDA path (FP64 filter, exact expansion arithmetic for the undecided ones) == oracle."""
import numpy as np
import pytest


def _pair_of_triangles(z0, z1, x=0.25):
    """a unit triangle in the plane z = 0 and a second triangle one of whose edges runs from height z0 to height z1 above
    the point (x, x) of the first"""
    V = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [x, x, z0], [x + 0.05, x + 0.05, z1], [x + 0.05, x, max(z0, z1) + 0.1]], float)
    F = np.array([[0, 1, 2], [3, 4, 5]], np.int32)
    return V, F


def cases(scenes):
    out = []
    for name, z0, z1, want in (("pierces", -0.5, 0.5, True), ("above", 0.2, 0.5, False), ("below", -0.5, -0.2, False),
                               ("grazing above (exact orientation needed)", 1e-300, 0.5, False),
                               ("grazing through (exact orientation needed)", -1e-300, 0.5, True),
                               ("end point exactly on the plane", 0.0, 0.5, True)):
        V, F = _pair_of_triangles(z0, z1)
        out.append((name, V, None, F, want))
    V0, V1, E, F, P = scenes.cloth_stack(3, 12)
    out.append(("stack (separated)", V0, E, F, False))
    W = V0.copy()
    per = V0.shape[0] // 3
    W[per + 70, 2] -= 1.5 * 0.5 * P["dhat"]  # one vertex of the middle sheet pushed through the sheet below
    out.append(("stack (one vertex pushed through)", W, E, F, True))
    V0, V1, E, F, P = scenes.cloth_on_sphere(48, 10, drape=True)  # fine enough that no flat cloth face dips below a sphere vertex
    out.append(("drape (separated)", V0, E, F, False))
    V0, V1, E, F, P = scenes.cloth_on_sphere(16, 10, drape=True)  # coarse: cloth faces sag 0.65 dhat, sphere vertices poke through
    out.append(("coarse drape (sphere vertices poke through the flat cloth faces)", V0, E, F, True))
    for seed in range(6):
        V0, V1, E, F, P = scenes.random_soup(12 + 6 * seed, seed=seed, scale=0.05 + 0.03 * seed)
        out.append(("soup %d" % seed, V0, E, F, None))
    return out


def run(api, scenes):
    res = {}
    for name, V, E, F, want in cases(scenes):
        if E is None:
            E = api.edges_from_faces(F)
        mesh = api.CollisionMesh(V, E, F)
        got = api.has_intersections(mesh, V)
        if want is not None:
            assert got == want, name
        res[name] = got
    return res


def test_has_intersections_oracle(oracle, scenes):
    res = run(oracle, scenes)
    soups = [v for k, v in res.items() if k.startswith("soup")]
    assert any(soups) and not all(soups)  # the random soups cover both answers


@pytest.mark.gpu
def test_has_intersections_gpu(cuda, oracle, scenes):
    assert run(cuda, scenes) == run(oracle, scenes)
    # a filtered mesh: the pierced sheets share a patch, so the crossing pair is not even a candidate (ipc.cpp:123)
    name, V, E, F, _ = [c for c in cases(scenes) if c[0].startswith("stack (one vertex")][0]
    for api in (cuda, oracle):
        mesh = api.CollisionMesh(V, E, F)
        mesh.can_collide = api.make_vertex_patches_filter(np.zeros(V.shape[0], np.int32))
        assert not api.has_intersections(mesh, V)
