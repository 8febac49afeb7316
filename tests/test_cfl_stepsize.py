"""Candidates::compute_noncandidate_conservative_stepsize / compute_cfl_stepsize (reference candidates.cpp:294-363,
SURVEY §8f rank 4).  CPU suite: the oracle against the definition evaluated in numpy; GPU suite: CUDA == oracle."""
import numpy as np
import pytest


def _stencil_vertices(cand, E, F):
    vs = [np.asarray(cand.vv_candidates).reshape(-1)]
    ev = np.asarray(cand.ev_candidates)
    vs += [ev[:, 1], E[ev[:, 0]].reshape(-1)] if len(ev) else []
    ee = np.asarray(cand.ee_candidates)
    vs += [E[ee].reshape(-1)] if len(ee) else []
    fv = np.asarray(cand.fv_candidates)
    vs += [fv[:, 1], F[fv[:, 0]].reshape(-1)] if len(fv) else []
    return np.unique(np.concatenate([np.asarray(v, np.int64) for v in vs]))


def _cases(scenes):
    V0, V1, E, F, P = scenes.cloth_on_sphere(24, 12)  # only the cap of the sheet has candidates
    yield "sphere", V0, V1, E, F, P["dhat"]
    V0, V1, E, F, P = scenes.cloth_stack(3, 12)
    yield "stack", V0, V1, E, F, P["dhat"]


def _ends(V0, V1, dhat):
    """end positions: a small step, the scene's step, and a rigid translation by 10 dhat (no impact, but far beyond the
    non-candidate bound: the full CCD has to run)"""
    return [V0 + 0.05 * (V1 - V0), V1, V0 + np.array([10 * dhat, 0, 0])]


def check(api, scenes, other=None):
    out = {}
    for name, V0, V1, E, F, dhat in _cases(scenes):
        mesh = api.CollisionMesh(V0, E, F)
        cand = api.Candidates()
        cand.build(mesh, V0, 0.5 * dhat)
        rng = np.random.default_rng(11)
        D = rng.normal(0, 3 * dhat, V0.shape)
        touched = _stencil_vertices(cand, E, F)
        assert 0 < len(touched) <= V0.shape[0]
        want = 0.5 * dhat / np.sqrt((D[touched] ** 2).sum(axis=1)).max()
        got = cand.compute_noncandidate_conservative_stepsize(mesh, D, dhat)
        assert abs(got - want) <= 1e-15 * want, name
        # CFL: a small step is bounded by min(alpha_C, alpha_F), a large one falls back to the full CCD
        res = []
        for W in _ends(V0, V1, dhat):
            cand.build(mesh, V0, 0.5 * dhat)
            a_c = cand.compute_collision_free_stepsize(mesh, V0, W)
            a_f = cand.compute_noncandidate_conservative_stepsize(mesh, W - V0, dhat)
            cfl = cand.compute_cfl_stepsize(mesh, V0, W, dhat)
            if a_f < 0.5 * a_c:
                mesh2 = api.CollisionMesh(V0, E, F)
                full = api.compute_collision_free_stepsize(mesh2, V0, W)
                assert abs(cfl - full) <= 1e-2 * full  # two runs of the same library (shared-bound pruning order)
            else:
                assert cfl == min(a_c, a_f)
            res.append((a_c, a_f, cfl))
        out[name] = (got, res)
    empty = api.Candidates()
    mesh = api.CollisionMesh(V0, E, F)
    empty.set(mesh)
    assert empty.compute_noncandidate_conservative_stepsize(mesh, V1 - V0, dhat) == 1.0  # candidates.cpp:301-303
    return out


def test_cfl_stepsizes_oracle(oracle, scenes):
    out = check(oracle, scenes)
    assert any(r[1] < 0.5 * r[0] for _, res in out.values() for r in res) and any(r[1] >= 0.5 * r[0] for _, res in out.values() for r in res)


@pytest.mark.gpu
def test_cfl_stepsizes_gpu(cuda, oracle, scenes):
    from ccd_tolerance import StepTolerance

    a, b = check(cuda, scenes), check(oracle, scenes)
    for (name, V0, V1, E, F, dhat) in _cases(scenes):
        assert a[name][0] == b[name][0]  # same arithmetic: bit-identical
        for W, (ra, rb) in zip(_ends(V0, V1, dhat), zip(a[name][1], b[name][1])):
            assert ra[1] == rb[1]
            tol = StepTolerance(oracle, V0, W, E, F).tolerance(min(ra[2], rb[2], 1.0))[0] if min(ra[2], rb[2]) < 1 else 0.0
            assert abs(ra[2] - rb[2]) <= max(tol, 1e-15), (name, ra, rb, tol)
