"""Derived tolerance of the Tight-Inclusion step size (north_star: "CCD step size within the configured tolerance and
never larger than the reference's by more than that tolerance").

The configured tolerance delta (TightInclusionCCD.tolerance, 1e-6) is a CO-DOMAIN tolerance — a distance.  Every valid
Tight-Inclusion search (the oracle's breadth-first one, the CUDA library's window-clipped depth-first / warp-cooperative
one) answers with the lower time bound t of a TERMINAL box [t, t + w] x U x V of the critical pair:

* the box was not rejected, so some point of it has |F|_inf <= ms + err + (co-domain width of the box), and a terminal
  box has co-domain width <= delta (or lies inside the eps-cube), i.e. the pair is within
  sqrt(3) (ms + err + delta) of contact distance somewhere in [t, t + w];
* no search answers later than the first time the pair reaches |F|_inf <= ms + err (that point's box is never
  rejected);
* the time width w of a terminal box is at most the domain tolerance of ccd.cu / oracle ccd.hpp `ti_tolerances`:
  tol_t = delta / (3 max_corner |relative displacement|_inf) <= delta / (sqrt(3) v) for a closing speed v (Euclidean,
  per unit of step) because v <= |relative displacement|_2 <= sqrt(3) |relative displacement|_inf.

Two valid answers therefore differ by at most the time the critical pair needs to close the distance
sqrt(3) (delta + err), plus one box width:

    |t_a - t_b|  <=  w + sqrt(3) (delta + err) / v,      w <= delta / (sqrt(3) v).

v is measured on the scene: the slope of the minimum distance over the swept candidates just before the oracle's step.
Because the pair that decides the minimum can differ between two searches (pairs have their own minimum separation
ms_i = min(0.2 d0_i, 1e-4), tight_inclusion_ccd.cpp:45-49), the bar used by the tests is TWICE that bound.
"""
import numpy as np

SQRT3 = 3.0 ** 0.5


class StepTolerance:
    def __init__(self, oracle, V0, V1, E, F, min_distance=0.0, tolerance=1e-6):
        self.o, self.V0, self.V1, self.md, self.delta = oracle, V0, V1, min_distance, tolerance
        self.mesh = oracle.CollisionMesh(V0, E, F)
        self.mesh2 = oracle.CollisionMesh(V0, E, F)  # is_step_collision_free rebuilds the candidates of its mesh
        self.cand = oracle.Candidates()
        self.cand.build(self.mesh, V0, V1, 0.5 * min_distance)
        lo, hi = np.minimum(V0.min(0), V1.min(0)), np.maximum(V0.max(0), V1.max(0))
        self.big = 10.0 * float(np.linalg.norm(hi - lo)) + 1.0
        mx = max(1.0, float(np.abs(V0).max()), float(np.abs(V1).max()))
        self.err = 7.549516567451064e-15 * mx ** 3  # the root finder's floating-point filter (ms > 0, vertex-face)

    def min_distance_at(self, t):
        """minimum distance over the swept candidates at V0 + t (V1 - V0)"""
        Vt = self.V0 + t * (self.V1 - self.V0)
        c = self.o.NormalCollisions()
        c.build(self.cand, self.mesh, Vt, self.big)
        return float(np.sqrt(c.compute_minimum_distance(self.mesh, Vt)))

    def closing_speed(self, t):
        h = min(0.02, 0.25 * t)
        if h <= 0:
            return 0.0
        return (self.min_distance_at(t - h) - self.min_distance_at(t)) / h

    def box_width(self, v):
        return self.delta / (SQRT3 * v)

    def tolerance(self, t_oracle):
        """(time tolerance, closing speed, box width) for a step size t_oracle < 1"""
        v = self.closing_speed(t_oracle)
        if not v > 0:
            return 1.0, v, 1.0  # no approach measured: the step size is not constrained by a closing pair
        w = self.box_width(v)
        return 2.0 * (w + SQRT3 * (self.delta + self.err) / v), v, w

    def is_step_collision_free(self, t):
        """the ORACLE's check of a step by t"""
        return self.o.is_step_collision_free(self.mesh2, self.V0, self.V0 + t * (self.V1 - self.V0), self.md)


def check_step(oracle, V0, V1, E, F, t_gpu, t_oracle, min_distance=0.0, tolerance=1e-6):
    """asserts the north_star bar for a Tight-Inclusion step size; returns the derived tolerance"""
    assert 0 <= t_gpu <= 1
    if t_oracle >= 1.0 and t_gpu >= 1.0:
        return 0.0
    T = StepTolerance(oracle, V0, V1, E, F, min_distance, tolerance)
    tol, v, w = T.tolerance(min(t_oracle, t_gpu))
    assert t_gpu <= t_oracle + tol, "step %.9g larger than the oracle's %.9g by more than the derived tolerance %.3e" % (t_gpu, t_oracle, tol)
    assert t_gpu >= t_oracle - tol, "step %.9g smaller than the oracle's %.9g by more than the derived tolerance %.3e" % (t_gpu, t_oracle, tol)
    if t_gpu > 0:
        # a step by the returned size, backed off by one box width, is collision free for the ORACLE's narrow phase
        assert T.is_step_collision_free(max(0.0, t_gpu - min(w, 0.5 * t_gpu))), "the oracle finds an impact inside the returned step"
    return tol
