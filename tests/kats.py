"""Known-answer / property vectors transcribed from the reference's own test-suite
(/root/reference/tests/src/tests, ipc-toolkit v1.6.0).  Each generator cites the test case it restates.
The reference draws its random inputs from Eigen's / Catch2's generators; here they come from a
seeded numpy Generator (the expectations are analytic, so any draw is a valid vector).  The very same
vectors are committed as tests/golden/reference_kats.npz by tests/golden/make_golden.py so that the
GPU box checks fixed numbers.

Everything returns plain numpy; nothing here imports the product or the oracle.
"""
import numpy as np

# distance-type enums (distance/distance_type.hpp:14-55)
P_E0, P_E1, P_E = 0, 1, 2
P_T0, P_T1, P_T2, PT_E0, PT_E1, PT_E2, P_T = range(7)
EA0_EB0, EA0_EB1, EA1_EB0, EA1_EB1, EA_EB0, EA_EB1, EA0_EB, EA1_EB, EA_EB = range(9)


def _unit(v):
    return v / np.linalg.norm(v)


def _edge_normal_2d(e0, e1):  # tests/src/tests/utils.hpp:93-99
    e = e1 - e0
    return _unit(np.array([-e[1], e[0]]))


# ---------------------------------------------------------------------------------------------------
def point_edge_type_cases(seed=11, n_random_edges=3):
    """distance/test_distance_type.cpp:14-56 (3D branch): returns (p, e0, e1, allowed dtypes)"""
    rng = np.random.default_rng(seed)
    out = []
    for alpha in np.arange(-1.0, 2.0, 0.1):
        for distance in np.arange(-10.0, 10.0, 1.0):
            for _ in range(n_random_edges):
                e0, e1 = rng.uniform(-1, 1, 3), rng.uniform(-1, 1, 3)
                n = _unit(np.cross(e1 - e0, [1.0, 0, 0]))
                p = ((e1 - e0) * alpha + e0) + distance * n
                if abs(alpha) < 1e-8:
                    ok = (P_E0, P_E)
                elif abs(alpha - 1) < 1e-8:
                    ok = (P_E1, P_E)
                elif alpha < 0:
                    ok = (P_E0,)
                elif alpha > 1:
                    ok = (P_E1,)
                else:
                    ok = (P_E,)
                out.append((p, e0, e1, ok))
    return out


def point_triangle_type_cases(seed=12, n_random=25):
    """distance/test_distance_type.cpp:95-226 + the GH issue vector :228-240: (p, t0, t1, t2, dtype)"""
    rng = np.random.default_rng(seed)
    t0, t1, t2 = np.array([-1.0, 0, 1]), np.array([1.0, 0, 1]), np.array([0.0, 0, -1])
    out = []

    def add(px, py, pz, expected):
        out.append((np.array([px, py, pz], float), t0, t1, t2, expected))

    for py in (-10, -1, -1e-12, 0, 1e-12, 1, 10):
        for pz in (0, -1 + 1e-12, 1 - 1e-12):  # closest to triangle
            add(0, py, pz, P_T)
        for _ in range(n_random):  # random closest to triangle
            margin = 1e-8
            bc = (rng.uniform(-1, 1, 2) + 1.0) / 2.0 * (1.0 - 2.0 * margin) + margin
            if bc.sum() >= 1:
                bc = 1 - bc
            b = np.array([bc[0], bc[1], 1 - bc.sum()])
            q = b[0] * t0 + b[1] * t1 + b[2] * t2
            add(q[0], py, q[2], P_T)
        for delta in (1e-8, 1e-4, 0.1, 11):
            add(t0[0] - delta, py, t0[2] + delta, P_T0)
            add(t1[0] + delta, py, t1[2] + delta, P_T1)
        for pz in (-1 - 1e-12, -1.1, -11):
            add(0, py, pz, P_T2)
        for (a, b, dt) in ((t0, t1, PT_E0), (t1, t2, PT_E1), (t2, t0, PT_E2)):
            perp = _edge_normal_2d(a[[0, 2]], b[[0, 2]])
            for alpha in (1e-4, 0.5, 1.0 - 1e-4):
                for scale in (1e-12, 1e-4, 1, 2, 11, 1000):
                    cp = (b - a) * alpha + a
                    add(cp[0] + scale * perp[0], py, cp[2] + scale * perp[1], dt)
            for _ in range(n_random):
                alpha, scale = rng.uniform(1e-8, 1 - 1e-8), rng.uniform(1e-12, 1e4)
                cp = (b - a) * alpha + a
                add(cp[0] + scale * perp[0], py, cp[2] + scale * perp[1], dt)
    out.append((np.array([0.488166, 0.0132623, 0.289055]), np.array([0.456476, 0.0526442, 0.260834]),
                np.array([0.609111, 0.0595969, 0.275928]), np.array([0.431262, 0.0508414, 0.255831]), PT_E0))
    return out


def _swap_table(swap_ea, swap_eb, swap_edges):
    """expected (ea0_eb0, ea1_eb0, ea_eb0) types after the swaps, distance/test_distance_type.cpp:293-330"""
    if not swap_edges:
        if swap_ea and swap_eb:
            return EA1_EB1, EA0_EB1, EA_EB1
        if swap_ea:
            return EA1_EB0, EA0_EB0, EA_EB0
        if swap_eb:
            return EA0_EB1, EA1_EB1, EA_EB1
        return EA0_EB0, EA1_EB0, EA_EB0
    if swap_ea and swap_eb:
        return EA1_EB1, EA1_EB0, EA1_EB
    if swap_ea:
        return EA0_EB1, EA0_EB0, EA0_EB
    if swap_eb:
        return EA1_EB0, EA1_EB1, EA1_EB
    return EA0_EB0, EA0_EB1, EA0_EB


def edge_edge_not_ea_eb_cases(seed=13, n_random_edges=2):
    """distance/test_distance_type.cpp:242-354 and distance/test_edge_edge.cpp:69-131:
    (ea0, ea1, eb0, eb1, allowed dtypes, expected squared distance s^2)"""
    rng = np.random.default_rng(seed)
    out = []
    sign = lambda x: -1 if x < 0 else 1
    for alpha0 in np.arange(-1.0, 2.0, 0.1):
        for s0 in np.arange(-10.0, 10.0, 1.0):
            if s0 == 0:
                continue
            for swap_ea in (False, True):
                for swap_eb in (False, True):
                    for swap_edges in (False, True):
                        alpha, s = alpha0, s0  # the reference mutates its generator copies per iteration
                        for _ in range(n_random_edges):
                            ea0, ea1 = rng.uniform(-1, 1, 3), rng.uniform(-1, 1, 3)
                            n = _unit(np.cross(ea1 - ea0, [1.0, 0, 0]))
                            eb0 = ((ea1 - ea0) * alpha + ea0) + s * n
                            if alpha < 0:
                                alpha = 0
                                n = eb0 - ea0
                                s = np.linalg.norm(n)
                                n = n / s
                            elif alpha > 1:
                                alpha = 1
                                n = eb0 - ea1
                                s = np.linalg.norm(n)
                                n = n / s
                            eb1 = ((ea1 - ea0) * alpha + ea0) + (s + sign(s) * 1) * n
                            a0, a1, b0, b1 = ea0, ea1, eb0, eb1
                            if swap_ea:
                                a0, a1 = a1, a0
                            if swap_eb:
                                b0, b1 = b1, b0
                            if swap_edges:
                                a0, a1, b0, b1 = b0, b1, a0, a1
                            t00, t10, t_0 = _swap_table(swap_ea, swap_eb, swap_edges)
                            if abs(alpha) <= 1e-15:
                                ok = (t00, t_0)
                            elif abs(alpha - 1) <= 1e-15:
                                ok = (t10, t_0)
                            elif alpha < 0:
                                ok = (t00,)
                            elif alpha > 1:
                                ok = (t10,)
                            else:
                                ok = (t_0,)
                            out.append((a0, a1, b0, b1, ok, s * s))
    return out


def edge_edge_ea_eb_cases(seed=14, n_random_edges=4):
    """distance/test_distance_type.cpp:356-402 and distance/test_edge_edge.cpp:133-176"""
    rng = np.random.default_rng(seed)
    out = []
    for alpha in rng.uniform(0.01, 0.99, 5):
        for beta in rng.uniform(0.01, 0.99, 5):
            for s in np.arange(-5.0, 5.0, 1.0):
                for swap_ea in (False, True):
                    for swap_eb in (False, True):
                        for swap_edges in (False, True):
                            for _ in range(n_random_edges):
                                ea1, eb1 = rng.uniform(-1, 1, 3), rng.uniform(-1, 1, 3)
                                ea0, eb0 = alpha / (alpha - 1) * ea1, beta / (beta - 1) * eb1
                                n = _unit(np.cross(ea1, eb1))
                                a0, a1, b0, b1 = ea0 - s / 2 * n, ea1 - s / 2 * n, eb0 + s / 2 * n, eb1 + s / 2 * n
                                if swap_ea:
                                    a0, a1 = a1, a0
                                if swap_eb:
                                    b0, b1 = b1, b0
                                if swap_edges:
                                    a0, a1, b0, b1 = b0, b1, a0, a1
                                out.append((a0, a1, b0, b1, (EA_EB,), s * s))
    return out


def edge_edge_parallel_cases(seed=15, n_random_edges=3):
    """distance/test_distance_type.cpp:404-439 (types) and distance/test_edge_edge.cpp:178-213 (distance):
    (ea0, ea1, eb0, eb1, allowed dtypes or None, s^2)"""
    rng = np.random.default_rng(seed)
    out = []
    for alpha in rng.uniform(0.01, 0.99, 10):
        for beta in rng.uniform(1.01, 1.99, 3):
            for s in rng.uniform(-5.0, 5.0, 10):
                for _ in range(n_random_edges):
                    ea0, ea1 = rng.uniform(-1, 1, 3), rng.uniform(-1, 1, 3)
                    ea = ea1 - ea0
                    n = _unit(np.cross(ea, [1.0, 0, 0]))
                    out.append((ea0, ea1, ea0 + alpha * ea + s * n, ea0 + beta * ea + s * n, (EA_EB0, EA1_EB), s * s))
                    eb0 = ea * alpha + ea0 + s * n  # test_edge_edge.cpp "distance parallel": eb = ea shifted
                    out.append((ea0, ea1, eb0, ea + eb0, None, s * s))
    return out


def edge_edge_coplanar_regression():
    """distance/test_distance_type.cpp:441-510: (ea0, ea1, eb0, eb1, required dtype or None, expected distance or None)"""
    out = [(np.array([-0.81818181276321411, 0.073941159961546266, 0.090909108519554152]),
            np.array([-0.81818181276321411, 0.073941161500775773, 0.272727280855178830]),
            np.array([-0.81818181276321411, 0.073941163540152718, 0.454545468091964780]),
            np.array([-0.81818181276321411, 0.073941167300585323, 0.636363625526428220]), EA1_EB0, None)]
    for L in (0.1, 0.5, 1.0, 2.0):
        for gap in (0.05, 0.1, 0.2, 0.5):
            for dy in (0.0, 1e-9, 1e-6):
                ea0, ea1 = np.array([0.0, 0, 0]), np.array([0.0, dy, L])
                eb0, eb1 = np.array([0.0, 2 * dy, L + gap]), np.array([0.0, 3 * dy, L + gap + L])
                out.append((ea0, ea1, eb0, eb1, None, float(np.sum((ea1 - eb0) ** 2))))
    return out


def point_triangle_distance_cases():
    """distance/test_point_triangle.cpp:31-110: (p, t0, t1, t2, closest point)"""
    t0, t1, t2 = np.array([-1.0, 0, 1]), np.array([1.0, 0, 1]), np.array([0.0, 0, -1])
    out = []
    for py in (-10, -1, -1e-12, 0, 1e-12, 1, 10):
        for pz in (0, -1 + 1e-12, -1, 1, 1 - 1e-12):
            out.append((np.array([0, py, pz], float), t0, t1, t2, np.array([0, 0, pz], float)))
        for px in (-1, -1 - 1e-12, -11):
            out.append((np.array([px, py, t0[2]], float), t0, t1, t2, t0))
        for px in (1, 1 + 1e-12, 11):
            out.append((np.array([px, py, t1[2]], float), t0, t1, t2, t1))
        for pz in (-1, -1 - 1e-12, -11):
            out.append((np.array([0, py, pz], float), t0, t1, t2, t2))
        for (a, b) in ((t0, t1), (t1, t2), (t2, t0)):
            perp = _edge_normal_2d(a[[0, 2]], b[[0, 2]])
            for alpha in (0.0, 1e-4, 0.5, 1.0 - 1e-4, 1.0):
                for scale in (0, 1e-12, 1e-4, 1, 2, 11, 1000):
                    cp = (b - a) * alpha + a
                    out.append((np.array([cp[0] + scale * perp[0], py, cp[2] + scale * perp[1]]), t0, t1, t2, cp))
    return out


def edge_edge_distance_grid():
    """distance/test_edge_edge.cpp:31-67: (e00, e01, e10, e11, expected squared distance)"""
    out = []
    for e0y in (-10, -1, -1e-4, 0, 1e-4, 1, 10):
        for shiftx in (-2, 0, 2):
            for shiftz in (-2, 0, 2):
                for dx in (-1, -0.5, 0, 0.5, 1):
                    for dz in (-1, -0.5, 0, 0.5, 1):
                        e0x, e0z = shiftx + dx, shiftz + dz
                        e00, e01 = np.array([-1 + e0x, e0y, e0z], float), np.array([1 + e0x, e0y, e0z], float)
                        e10, e11 = np.array([0.0, 0, -1]), np.array([0.0, 0, 1])
                        c0 = e00 if shiftx > 1 else (e01 if shiftx < -1 else np.array([0, e0y, e0z], float))
                        c1 = e11 if shiftz > 1 else (e10 if shiftz < -1 else np.array([0, 0, e0z], float))
                        out.append((e00, e01, e10, e11, float(np.sum((c0 - c1) ** 2))))
    return out


def edge_edge_degenerate_cases():
    """distance/test_edge_edge.cpp:215-265: rotating degenerate edges and non-overlapping collinear edges"""
    out = []
    for e0y in (-10, -1, -1e-4, 0, 1e-4, 1, 10):
        for th in (-2, -1.5, -1, -0.123124, 0, 0.2342352, 0.5, 1, 1.5, 2, 50, 51):
            c, s = np.cos(th * np.pi), np.sin(th * np.pi)
            R = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])  # AngleAxis(theta, UnitY)
            e00, e01 = np.array([-1.0, e0y, 0]), np.array([1.0, e0y, 0])
            e10, e11 = np.array([0.0, 0, -1]), np.array([0.0, 0, 1])
            out.append((R @ e00, R @ e01, e10, e11, e0y * e0y))
            out.append((e00, e01, R @ e10, R @ e11, e0y * e0y))
        for gap in (0, 0.01, 0.1, 1):
            e00, e01 = np.array([gap, e0y, 0], float), np.array([1, e0y, 0], float)
            e10, e11 = np.array([-1.0, 0, 0]), np.array([-gap, 0, 0], float)
            exp = float(np.sum((np.array([gap, e0y, 0]) - np.array([-gap, 0, 0])) ** 2))
            for (a, b, c_, d) in ((e00, e01, e10, e11), (e01, e00, e10, e11), (e00, e01, e11, e10), (e01, e00, e11, e10)):
                out.append((a, b, c_, d, exp))
    return out


# ---------------------------------------------------------------------------------------------------
EPS_F = float(np.finfo(np.float32).eps)  # ccd/test_edge_edge_ccd.cpp:10, test_point_triangle_ccd.cpp:11


def edge_edge_ccd_cases():
    """ccd/test_edge_edge_ccd.cpp:20-165: dicts with t0/t1 stencils (ea0, ea1, eb0, eb1), expectation, tol, max_iter, tmax"""
    out = []

    def add(name, t0, t1, expected, conservative=True, tol=1e-6, max_iter=10_000_000, tmax=1.0):
        out.append(dict(name=name, t0=np.array(t0, float), t1=np.array(t1, float), expected=bool(expected), conservative=conservative,
                        tol=tol, max_iter=max_iter, tmax=tmax))

    for uy in (-1.0, 0.0, 1 - EPS_F, 1.0, 1 + EPS_F, 2.0):
        for e1x in (-1 - EPS_F, -1, -1 + EPS_F, -0.5, 0, 0.5, 1 - EPS_F, 1, 1 + EPS_F):
            t0 = [[-1, -1, 0], [1, -1, 0], [e1x, 1, -1], [e1x, 1, 1]]
            for (u0, u1, exp) in (((0, uy, 0), (0, -uy, 0), uy >= 1.0 and -1 <= e1x <= 1),
                                  ((0, 2 * uy, 0), (0, 0, 0), uy >= 2.0 and -1 <= e1x <= 1)):
                t1 = [np.add(t0[0], u0), np.add(t0[1], u0), np.add(t0[2], u1), np.add(t0[3], u1)]
                add("general", t0, t1, exp)
    add("double root 1", [[-3.0022200, 0.2362580, 0.0165247], [-3.2347850, 0.8312380, -0.1151003], [-3.0319900, 0.3148750, 0],
                          [-2.8548800, 0.0900349, 0]],
        [[-2.8995600, 0.0345838, 0.0638580], [-3.1716930, 0.6104858, -0.0713340], [-3.0319900, 0.3148750, 0], [-2.8548800, 0.0900349, 0]],
        True)
    for t in (0.5, 0.8, 0.88, 0.9, 1.0):
        a0, a1 = np.array([0.0, 0, 1]), np.array([0.0, 1, 1])
        add("double root 2", [a0, a1, [0.1, 0.2, 2], [0.1, 0.2, -1]],
            [(np.array([1.0, 1, 0]) - a0) * t + a0, (np.array([0.0, 0, 0]) - a1) * t + a1, [0.1, 0.2, 2], [0.1, 0.2, -1]], True)
    for tol in (1e-6, 1e-5, 1e-4, 1e-3, 1e-2, 1e-1):
        add("slow case 1", [[1, 0.50803125, 2.10835646075301e-18], [-2.38233935445388e-18, 0.50803125, 1],
                            [-4.99999999958867e-07, 0.5, 0], [-4.99999999958867e-07, 0.5, 1]],
            [[1, 0.47124375, 4.11078309465837e-18], [-2.8526707189104e-18, 0.47124375, 1], [-4.99999999958867e-07, 0.5, 0],
             [-4.99999999958867e-07, 0.5, 1]], True, conservative=False, tol=tol)
    add("slow case 2", [[1.00002232466453, 0.500004786049044, -2.06727783590977e-05],
                        [1.64687846177844e-05, 0.499996645067319, 1.63939999009028e-05], [1, 0.5, 0], [0, 0.5, 0]],
        [[1.00294282700155, 0.498652627047143, 0.003626320742036], [-0.00219276550735626, 0.500871179186644, -0.00315828804921928],
         [1, 0.5, 0], [0, 0.5, 0]], True, conservative=False, max_iter=1_000_000, tmax=2.8076171875e-03)
    for d0 in (1e-2, 1e-4, 1e-6, 1e-8):
        t0 = np.array([[-1.0, -d0 / 2, 0], [1.0, -d0 / 2, 0], [0, d0 / 2, -1.0], [0, d0 / 2, 1.0]])
        dy, sc = d0, 1e-3
        t1 = t0 + np.array([[-sc, dy, -sc], [sc, dy, sc], [sc, -dy, sc], [-sc, -dy, -sc]])
        add("adversarial accd", t0, t1, dy >= d0 / 2, conservative=False)
    return out


def point_triangle_ccd_cases():
    """ccd/test_point_triangle_ccd.cpp:18-123 (NDEBUG branch): stencils (p, t0, t1, t2)"""
    out = []

    def add(name, t0, t1, expected, conservative=True):
        out.append(dict(name=name, t0=np.array(t0, float), t1=np.array(t1, float), expected=bool(expected), conservative=conservative,
                        tol=1e-6, max_iter=10_000_000, tmax=1.0))

    vals = (-1.0, 0.0, 0.5 - EPS_F, 0.5, 0.5 + EPS_F, 1.0, 2.0)
    for v0z in (0.0, -1.0):
        for g in vals:
            u0y = -g
            for u0z in (-EPS_F, 0.0, EPS_F):
                for u1y in vals:
                    t0 = np.array([[0, 1, v0z], [-1, 0, 1], [1, 0, 1], [0, 0, -1]], float)
                    t1 = t0 + np.array([[0, u0y, u0z], [0, u1y, 0], [0, u1y, 0], [0, u1y, 0]])
                    add("general", t0, t1, (-u0y + u1y >= 1) and (v0z + u0z >= -1))
    for qy in (-EPS_F, 0, EPS_F):
        add("zhongshi", [[0, qy, 0], [0, 0, 0], [0, 1, 0], [1, 0, 0]], [[0, qy, 0], [0, 0, 1], [0, 1, 1], [1, 0, 1]], qy >= 0)
    add("bolun", [[0.1, 0.1, 0.1], [0, 0, 1], [1, 0, 1], [0, 1, 1]], [[0.1, 0.1, 0.1], [0, 0, 0], [0, 1, 0], [1, 0, 0]], True)
    add("no zero toi", [[0.0133653, 0.100651, -0.0215935], [0.0100485, 0.0950896, -0.0171013], [0.0130388, 0.100666, -0.0218112],
                        [0.015413, 0.100554, -0.0202265]],
        [[0.0133652999767858, 0.099670000268615, -0.0215934999996444], [0.0100484999799995, 0.0941086002577558, -0.0171012999972189],
         [0.0130387999724314, 0.0996850002629403, -0.0218111999936902], [0.0154129999740718, 0.0995730002646605, -0.020226499996014]],
        False, conservative=False)
    return out


def point_edge_ccd_cases():
    """ccd/test_point_edge_ccd.cpp:41-135, 3D embedding (z = 0) of the 2D sections: stencils (p, e0, e1); toi or None"""
    S = []

    def add(name, p0, a0, b0, p1, a1, b1, expected, toi=None):
        z = lambda v: [v[0], v[1], 0.0]
        S.append(dict(name=name, t0=np.array([z(p0), z(a0), z(b0), [0, 0, 0]], float), t1=np.array([z(p1), z(a1), z(b1), [0, 0, 0]], float),
                      expected=expected, toi=toi))

    add("degenerate before impact", (0, 1), (-1, 0), (1, 0), (0, -1), (3, 0), (-3, 0), True, 0.5)
    add("degenerate after impact", (0, 1), (-1, 0), (1, 0), (0, -1), (0.5, 0), (-0.5, 0), True, 0.5)
    add("edge right point left", (-1, 0), (1, -1), (1, 1), (1, 0), (-1, -1), (-1, 1), True, 0.5)
    add("point on edge line", (0, 0), (0, 1), (0, 2), (0, 2), (0, 1), (0, 2), True, 0.5)
    add("parallel", (0, 1), (1, 0), (1, 2), (0, 2), (1, 1), (1, 3), False)
    add("stretching e0=[1,2]", (0, 0), (1, 1), (1, -1), (1, 0), (1, 2), (1, -2), True, 1.0)
    add("stretching e0=[0,2]", (0, 0), (1, 1), (1, -1), (1, 0), (1, 2), (1, -2), True, 1.0)  # :113-123 swaps p and e0: same set
    add("point-point", (1.11111, 0.5), (1, 0.5), (1, 0.75), (0.888889, 0.5), (1, 0.5), (1, 0.75), True, 0.5)
    return S


def point_point_ccd_cases():
    """ccd/test_point_point_ccd.cpp:10-38: p0 (0,0,0)->(1,1,1), p1 (1,1,0)->(0,0,1): toi 0.5 +- 1e-3"""
    return dict(t0=np.array([[0.0, 0, 0], [1.0, 1, 0], [0, 0, 0], [0, 0, 0]]), t1=np.array([[1.0, 1, 1], [0.0, 0, 1], [0, 0, 0], [0, 0, 0]]),
                min_distances=(0, 1e-6, 1e-4, 1e-2), toi=0.5, margin=1e-3)


# ---------------------------------------------------------------------------------------------------
def barrier_potential_scenes():
    """potential/test_barrier_potential.cpp:133-212 (3D sections): name -> (V, E, F, dhat)"""
    dhat = 1e-3
    fv_V = np.array([[0, 1e-4, 0], [-1, 0, 0], [1e-4, 0, -1], [1e-4, 0, 1], [1, 0, 0]], float)
    fv_F = np.array([[1, 2, 3], [2, 3, 4]], np.int32)
    ee_V = np.array([[0, 1e-4, -1], [0, 1e-4, 1], [-1e-4, 0, 0], [-1, 0, 0], [1, 0, 0]], float)
    ee_E = np.array([[0, 1], [3, 2], [2, 4]], np.int32)
    par_V = np.array([[-0.5, 1e-5, -1e-3], [0.5, 1e-5, 1e-3], [-1, -1e-5, 0], [0, -1e-5, 0], [1, -1e-5, 0]], float)
    par_E = np.array([[0, 1], [2, 3], [3, 4]], np.int32)
    return {"3D Face-Vertex": (fv_V, None, fv_F, dhat), "3D Edge-Edge": (ee_V, ee_E, None, dhat),
            "3D Edge-Edge Parallel": (par_V, par_E, None, dhat)}


def readme_quick_start():
    """README.md:51-88: two parallel triangles, gap 0.5*dhat"""
    dhat = 1e-3
    gap = 0.5 * dhat
    V = np.array([[0, 0, 0], [1, 0, 0], [0.5, 1, 0], [0, 0, gap], [1, 0, gap], [0.5, 1, gap]], float)
    E = np.array([[0, 1], [1, 2], [2, 0], [3, 4], [4, 5], [5, 3]], np.int32)
    F = np.array([[0, 1, 2], [3, 4, 5]], np.int32)
    V1 = V.copy()
    V1[3:, 2] -= 2 * gap
    return V, E, F, V1, dhat
