"""Synthetic scenes for the configurations named in BASELINE.json.

All generators return ``(V0, V1, E, F, params)``: rest == V0, V1 = end-of-step
positions for CCD, int32 edges / faces, and a dict with ``dhat`` etc.  Vertices
are jittered (seeded) so that no distance-type parameter lands exactly on a
region boundary (SURVEY §7 hard part 2).
"""
import numpy as np


def _edges(F):
    e = np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]]).astype(np.int64)
    e.sort(axis=1)
    return np.unique(e, axis=0).astype(np.int32)


def grid_sheet(nx, ny, size=1.0):
    """(nx+1)*(ny+1) vertices on [-size/2, size/2]^2, 2*nx*ny triangles"""
    xs = np.linspace(-0.5 * size, 0.5 * size, nx + 1)
    ys = np.linspace(-0.5 * size, 0.5 * size, ny + 1)
    X, Y = np.meshgrid(xs, ys, indexing="ij")
    V = np.stack([X.ravel(), Y.ravel(), np.zeros(X.size)], axis=1)
    idx = np.arange((nx + 1) * (ny + 1)).reshape(nx + 1, ny + 1)
    a, b, c, d = idx[:-1, :-1].ravel(), idx[1:, :-1].ravel(), idx[1:, 1:].ravel(), idx[:-1, 1:].ravel()
    F = np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d], 1)]).astype(np.int32)
    return V, F


def uv_sphere(n_lat, n_lon, radius=1.0):
    """closed latitude / longitude sphere with two pole vertices: 2*n_lon*(n_lat-1) triangles"""
    th = np.linspace(0, np.pi, n_lat + 1)[1:-1]
    ph = np.linspace(0, 2 * np.pi, n_lon, endpoint=False)
    T, P = np.meshgrid(th, ph, indexing="ij")
    ring = np.stack([np.sin(T) * np.cos(P), np.sin(T) * np.sin(P), np.cos(T)], axis=-1).reshape(-1, 3)
    V = np.concatenate([[[0, 0, 1.0]], ring, [[0, 0, -1.0]]]) * radius
    F = []
    r = lambda i, j: 1 + i * n_lon + (j % n_lon)
    south = V.shape[0] - 1
    for j in range(n_lon):
        F.append([0, r(0, j), r(0, j + 1)])
        F.append([south, r(n_lat - 2, j + 1), r(n_lat - 2, j)])
    for i in range(n_lat - 2):
        for j in range(n_lon):
            F.append([r(i, j), r(i + 1, j), r(i + 1, j + 1)])
            F.append([r(i, j), r(i + 1, j + 1), r(i, j + 1)])
    return V, np.asarray(F, dtype=np.int32)


def _merge(parts):
    Vs, Fs, off = [], [], 0
    for V, F in parts:
        Vs.append(V)
        Fs.append(F + off)
        off += V.shape[0]
    return np.concatenate(Vs), np.concatenate(Fs).astype(np.int32)


def cloth_on_sphere(n_cloth=64, sphere_res=32, dhat=1e-3, seed=1, drape=False, offset=None):
    """C1 (flat sheet 0.6*dhat above the north pole) / C2 (drape=True: the sheet follows a sphere of
    radius 1 + offset*dhat so the whole sheet is in contact)"""
    rng = np.random.default_rng(seed)
    h = 1.0 / n_cloth
    Vc, Fc = grid_sheet(n_cloth, n_cloth, 1.0)
    Vc[:, :2] += rng.uniform(-0.05, 0.05, (Vc.shape[0], 2)) * h
    if drape:
        rho = 1.0 + (0.5 if offset is None else offset) * dhat
        Vc[:, 2] = np.sqrt(rho * rho - Vc[:, 0] ** 2 - Vc[:, 1] ** 2)
    else:
        Vc[:, 2] = 1.0 + (0.6 if offset is None else offset) * dhat
    Vc += rng.uniform(-1e-6, 1e-6, Vc.shape) * h
    Vs, Fs = uv_sphere(sphere_res, sphere_res)
    V0, F = _merge([(Vc, Fc), (Vs, Fs)])
    V1 = V0.copy()
    V1[: Vc.shape[0], 2] -= 2 * dhat  # the cloth moves down through the contact gap
    return V0, V1, _edges(F), F, {"dhat": dhat, "n_cloth_vertices": Vc.shape[0]}


def cloth_stack(layers=8, n=250, dhat=1e-3, gap=0.5, seed=3, h=None):
    """C3 / C5: `layers` stacked n x n cloth sheets, gap*dhat apart, each rotated by a distinct small
    angle so that edges cross (dense edge-edge contact).  `h` fixes the cell size (default 1/n, i.e. a
    unit sheet); a smaller n with the same h is a crop of the same workload (bounded CPU samples)."""
    rng = np.random.default_rng(seed)
    h = 1.0 / n if h is None else h
    parts = []
    for k in range(layers):
        V, F = grid_sheet(n, n, n * h)
        ang = 0.05 + 0.11 * k
        c, s = np.cos(ang), np.sin(ang)
        V[:, :2] = V[:, :2] @ np.array([[c, -s], [s, c]]).T
        V[:, 2] = k * gap * dhat
        V[:, :2] += rng.uniform(-1e-3, 1e-3, (V.shape[0], 2)) * h
        V[:, 2] += rng.uniform(-1e-2, 1e-2, V.shape[0]) * dhat
        parts.append((V, F))
    V0, F = _merge(parts)
    V1 = V0.copy()
    nv = parts[0][0].shape[0]
    for k in range(layers):  # squeeze the stack past contact: neighbouring layers meet at t ~ 0.77
        V1[k * nv:(k + 1) * nv, 2] -= (k - 0.5 * (layers - 1)) * 1.3 * gap * dhat
    return V0, V1, _edges(F), F, {"dhat": dhat, "layers": layers}


def perturbed_sheets(layers=4, n=64, spacing=2.0, disp=5.0, seed=4):
    """C4-style broad-phase + CCD stress: stacked sheets `spacing`*h apart whose vertices are displaced
    by U(-1,1)*disp*h during the step (large swept boxes, high candidate count)"""
    rng = np.random.default_rng(seed)
    h = 1.0 / n
    parts = []
    for k in range(layers):
        V, F = grid_sheet(n, n, 1.0)
        V[:, 2] = k * spacing * h
        V += rng.uniform(-0.1, 0.1, V.shape) * h
        parts.append((V, F))
    V0, F = _merge(parts)
    V1 = V0 + rng.uniform(-1, 1, V0.shape) * disp * h
    return V0, V1, _edges(F), F, {"dhat": 0.6 * spacing * h}


def dense_sheet(n=16, dhat_cells=3.0, seed=6):
    """one jittered n x n sheet whose dhat spans `dhat_cells` cells: every vertex is in contact with
    hundreds of primitives of its own sheet (stress for per-vertex work such as Hessian columns with
    thousands of row blocks)"""
    rng = np.random.default_rng(seed)
    h = 1.0 / n
    V, F = grid_sheet(n, n, 1.0)
    V[:, :2] += rng.uniform(-0.2, 0.2, (V.shape[0], 2)) * h
    V[:, 2] = rng.uniform(-0.3, 0.3, V.shape[0]) * h
    V1 = V + rng.normal(0, 0.3, V.shape) * h
    return V, V1, _edges(F), F, {"dhat": dhat_cells * h}


def random_soup(n_tris=200, seed=0, scale=0.15):
    """small random triangle soup (every triangle has its own 3 vertices) — dense, irregular contact"""
    rng = np.random.default_rng(seed)
    c = rng.uniform(0, 1, (n_tris, 1, 3))
    V0 = (c + rng.normal(0, scale, (n_tris, 3, 3))).reshape(-1, 3)
    F = np.arange(3 * n_tris, dtype=np.int32).reshape(-1, 3)
    V1 = V0 + rng.normal(0, 0.5 * scale, V0.shape)
    return V0, V1, _edges(F), F, {"dhat": 0.05}
