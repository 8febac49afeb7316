"""Host-side mirror of the reference's Python interface (``ipctk``) for the
per-step contact path, on top of the C ABI (include/ipcb200.h).

Names, argument meaning and error behaviour follow the reference's pybind11
module (python/src/**): ``CollisionMesh``, ``Candidates``, ``NormalCollisions``,
``BarrierPotential``, ``compute_collision_free_stepsize``,
``TightInclusionCCD`` / ``AdditiveCCD``, ``PSDProjectionMethod``.

``make_api(lib)`` returns a namespace bound to one loaded library: the product
(``ipcb_``; CUDA, no CPU fallback) or — from tests only — the CPU oracle.

State model: a ``CollisionMesh`` owns the library context (device mirrors of
the mesh, the broad phase, and ONE resident candidate set and ONE resident
collision set).  ``Candidates`` / ``NormalCollisions`` objects are handles to
that resident state plus a lazily fetched host copy (the
public host containers are the compatibility path; the timed path never
materialises them).  Building a newer set on the same mesh invalidates older
handles (using them raises ``RuntimeError``).
"""
import ctypes as C
import enum
import types
import weakref

import numpy as np

try:  # loaded as a package module, or stand-alone by path (the test-only checker does that)
    from . import _abi
except ImportError:  # pragma: no cover
    import _abi

VV, EV, EE, FV, EF, FF = range(6)


class PSDProjectionMethod(enum.IntEnum):  # utils/eigen_ext.hpp:202-206
    NONE = 0
    CLAMP = 1
    ABS = 2


class TightInclusionCCD:  # ccd/tight_inclusion_ccd.hpp:11-19
    DEFAULT_TOLERANCE = 1e-6
    DEFAULT_MAX_ITERATIONS = 10_000_000
    DEFAULT_CONSERVATIVE_RESCALING = 0.8
    SMALL_TOI = 1e-6

    def __init__(self, tolerance=DEFAULT_TOLERANCE, max_iterations=DEFAULT_MAX_ITERATIONS,
                 conservative_rescaling=DEFAULT_CONSERVATIVE_RESCALING):
        self.tolerance = tolerance
        self.max_iterations = max_iterations
        self.conservative_rescaling = conservative_rescaling

    def _params(self):
        return _abi.CcdParams(0, self.tolerance, self.max_iterations, self.conservative_rescaling)


class AdditiveCCD:  # ccd/additive_ccd.hpp:21-26
    DEFAULT_MAX_ITERATIONS = 10_000_000
    DEFAULT_CONSERVATIVE_RESCALING = 0.9

    def __init__(self, max_iterations=DEFAULT_MAX_ITERATIONS, conservative_rescaling=DEFAULT_CONSERVATIVE_RESCALING):
        self.max_iterations = max_iterations
        self.conservative_rescaling = conservative_rescaling

    def _params(self):
        return _abi.CcdParams(1, 0.0, self.max_iterations, self.conservative_rescaling)


class CollisionFilter:
    """ipc::CollisionFilter (collision_filter.hpp:30-111) restricted to what is DATA: the intersection of a vertex-patch
    filter and a static-obstacle filter (the factories of collision_filter.hpp:113-143).  `a & b` composes two of them;
    unions, negations and arbitrary callables cannot cross the C ABI (the C++ adapter post-filters those on the host)."""

    def __init__(self, patch_ids=None, n_dynamic=None):
        self.patch_ids = None if patch_ids is None else np.ascontiguousarray(patch_ids, dtype=np.int32).reshape(-1)
        self.n_dynamic = None if n_dynamic is None else int(n_dynamic)

    def __call__(self, vi, vj):
        ok = True
        if self.patch_ids is not None:
            ok = ok and self.patch_ids[vi] != self.patch_ids[vj]
        if self.n_dynamic is not None:
            ok = ok and (vi < self.n_dynamic or vj < self.n_dynamic)
        return bool(ok)

    def __and__(self, other):
        if self.patch_ids is not None and other.patch_ids is not None:
            raise NotImplementedError("the intersection of two patch filters is not a patch filter")
        n = [x for x in (self.n_dynamic, other.n_dynamic) if x is not None]
        return CollisionFilter(self.patch_ids if self.patch_ids is not None else other.patch_ids, min(n) if n else None)


def make_vertex_patches_filter(patch_ids):  # collision_filter.hpp:113-118
    return CollisionFilter(patch_ids=patch_ids)


def make_static_obstacle_filter(n_dynamic):  # collision_filter.hpp:120-131
    return CollisionFilter(n_dynamic=n_dynamic)


def make_connected_components_filter(faces, num_vertices=None):  # collision_filter.cpp:9-18
    import scipy.sparse as sp
    from scipy.sparse.csgraph import connected_components

    f = np.asarray(faces, dtype=np.int64).reshape(-1, 3)
    n = int(f.max()) + 1 if num_vertices is None else int(num_vertices)
    i = np.concatenate([f[:, 0], f[:, 1], f[:, 2]])
    j = np.concatenate([f[:, 1], f[:, 2], f[:, 0]])
    A = sp.coo_matrix((np.ones(i.size), (i, j)), shape=(n, n))
    return CollisionFilter(patch_ids=connected_components(A, directed=False)[1])


def _f64(a):
    """column-major float64 view/copy (Eigen::MatrixXd layout) and its leading dimension"""
    a = np.asfortranarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.c_void_p), (a.shape[0] if a.ndim == 2 else a.size)


def _i32(a, cols):
    a = np.asarray(a, dtype=np.int32)
    if a.size == 0:
        a = a.reshape(0, cols)
    a = np.asfortranarray(a)
    return a, a.ctypes.data_as(C.c_void_p), max(a.shape[0], 1)


def edges_from_faces(faces):
    """unique undirected edges of a triangle list (the reference uses igl::edges)"""
    f = np.asarray(faces, dtype=np.int64).reshape(-1, 3)
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
    e.sort(axis=1)
    return np.unique(e, axis=0).astype(np.int32)


def make_api(lib):
    ns = types.SimpleNamespace()
    ns.lib = lib
    ns.PSDProjectionMethod = PSDProjectionMethod
    ns.TightInclusionCCD = TightInclusionCCD
    ns.AdditiveCCD = AdditiveCCD
    ns.edges_from_faces = edges_from_faces
    ns.CollisionFilter = CollisionFilter
    ns.make_vertex_patches_filter = make_vertex_patches_filter
    ns.make_static_obstacle_filter = make_static_obstacle_filter
    ns.make_connected_components_filter = make_connected_components_filter

    class CollisionMesh:
        """ipc::CollisionMesh(rest_positions, edges, faces) — collision_mesh.cpp:15-127"""

        def __init__(self, rest_positions, edges=None, faces=None, device=0):
            rest = np.asarray(rest_positions, dtype=np.float64)
            if rest.ndim != 2 or rest.shape[1] != 3:
                raise ValueError("only 3D meshes are supported on this path (rest_positions must be N x 3)")
            self.rest_positions = rest
            self.edges = np.zeros((0, 2), np.int32) if edges is None else np.asarray(edges, np.int32).reshape(-1, 2)
            self.faces = np.zeros((0, 3), np.int32) if faces is None else np.asarray(faces, np.int32).reshape(-1, 3)
            self._ctx = C.c_void_p()
            lib.check(lib.ctx_create(device, C.byref(self._ctx)))
            r, rp, rld = _f64(rest)
            e, ep, eld = _i32(self.edges, 2)
            f, fp, fld = _i32(self.faces, 3)
            lib.check(lib.mesh_set(self._ctx, rest.shape[0], rp, rld, e.shape[0], ep, eld, f.shape[0], fp, fld))
            self._cand_gen = 0
            self._resident = None  # weak reference to the NormalCollisions whose records are the context's resident set
            self._can_collide = CollisionFilter()

        @property
        def can_collide(self):
            """CollisionMesh::can_collide (collision_mesh.hpp:338): a CollisionFilter descriptor"""
            return self._can_collide

        @can_collide.setter
        def can_collide(self, f):
            if not isinstance(f, CollisionFilter):
                raise TypeError("can_collide must be a CollisionFilter descriptor (patch labels and / or n_dynamic)")
            ids = f.patch_ids
            if ids is not None and ids.size != self.num_vertices():
                raise ValueError("patch_ids must hold one label per vertex")
            lib.check(lib.mesh_set_collision_filter(self._ctx, None if ids is None else ids.ctypes.data_as(C.c_void_p),
                                                    -1 if f.n_dynamic is None else f.n_dynamic))
            self._can_collide = f
            self._cand_gen += 1  # resident candidates of the old filter are stale

        def close(self):
            """destroy the library context now (device buffers, streams); the object is unusable afterwards"""
            if getattr(self, "_ctx", None):
                lib.ctx_destroy(self._ctx)
                self._ctx = None

        def __del__(self):
            self.close()

        def num_vertices(self):
            return self.rest_positions.shape[0]

        def num_edges(self):
            return self.edges.shape[0]

        def num_faces(self):
            return self.faces.shape[0]

        def dim(self):
            return 3

        def num_codim_vertices(self):
            n = C.c_int32()
            lib.check(lib.mesh_num_codim_vertices(self._ctx, C.byref(n)))
            return n.value

        def num_codim_edges(self):
            n = C.c_int32()
            lib.check(lib.mesh_num_codim_edges(self._ctx, C.byref(n)))
            return n.value

        def faces_to_edges(self):
            out = np.zeros((self.num_faces(), 3), np.int32, order="F")
            lib.check(lib.mesh_faces_to_edges(self._ctx, out.ctypes.data_as(C.c_void_p)))
            return out

        def vertex_areas(self):
            return self._areas()[0]

        def edge_areas(self):
            return self._areas()[1]

        def set_broad_phase_method(self, method):
            """"lbvh" (default) or "sap": the CUDA broad phase behind every later build on this mesh (product only)"""
            m = {"lbvh": 0, "sap": 1}[method] if isinstance(method, str) else int(method)
            lib.check(lib.ctx_set_broad_phase_method(self._ctx, m))
            self._cand_gen += 1

        # ---- sharding of the potential over ranks holding the same collision set (include/ipcb200.h)
        def set_collision_range(self, rank, world):
            lib.check(lib.ctx_set_collision_range(self._ctx, rank, world))

        def set_row_block(self, v_begin=0, v_end=-1):
            lib.check(lib.ctx_set_row_block(self._ctx, v_begin, v_end))

        def balanced_row_blocks(self, world):
            bounds = np.zeros(world + 1, np.int32)
            lib.check(lib.hessian_balanced_row_blocks(self._ctx, world, bounds.ctypes.data_as(C.c_void_p)))
            return bounds

        def _areas(self):
            va = np.zeros(self.num_vertices())
            ea = np.zeros(self.num_edges())
            lib.check(lib.mesh_areas(self._ctx, va.ctypes.data_as(C.c_void_p), ea.ctypes.data_as(C.c_void_p)))
            return va, ea

    class BroadPhase:
        """The CUDA LBVH behind ipc::BroadPhase (broad_phase/broad_phase.hpp:19-133).

        build(...) then detect_*_candidates() -> (n, 2) int32 arrays sorted
        lexicographically (unordered kinds as (min, max))."""

        FLOAT, DOUBLE = 0, 1

        def __init__(self, mesh, boxes=0):
            self.mesh = mesh
            self.boxes = boxes

        def name(self):
            return "CudaLBVH" if lib.has_device_api else "OracleLBVH"

        def build(self, vertices_t0, vertices_t1=None, inflation_radius=0.0):
            v0, p0, ld = _f64(vertices_t0)
            if vertices_t1 is None:
                lib.check(lib.broad_build_static(self.mesh._ctx, p0, ld, inflation_radius, self.boxes))
            else:
                v1, p1, _ = _f64(vertices_t1)
                lib.check(lib.broad_build_swept(self.mesh._ctx, p0, p1, ld, inflation_radius, self.boxes))

        def vertex_boxes(self):
            out = np.zeros((self.mesh.num_vertices(), 6), np.float32 if self.boxes == 0 else np.float64)
            lib.check(lib.broad_vertex_boxes(self.mesh._ctx, out.ctypes.data_as(C.c_void_p)))
            return out

        def _detect(self, kind):
            n = C.c_int64()
            lib.check(lib.broad_detect(self.mesh._ctx, kind, C.byref(n)))
            out = np.zeros((n.value, 2), np.int32)
            if n.value:
                lib.check(lib.broad_fetch(self.mesh._ctx, kind, out.ctypes.data_as(C.c_void_p)))
            return out

        def detect_vertex_vertex_candidates(self):
            return self._detect(VV)

        def detect_edge_vertex_candidates(self):
            return self._detect(EV)

        def detect_edge_edge_candidates(self):
            return self._detect(EE)

        def detect_face_vertex_candidates(self):
            return self._detect(FV)

        def detect_edge_face_candidates(self):
            return self._detect(EF)

        def detect_face_face_candidates(self):
            return self._detect(FF)

    class Candidates:
        """ipc::Candidates — candidates/candidates.cpp:43-292"""

        def __init__(self):
            self.mesh = None
            self._gen = -1
            self._counts = [0, 0, 0, 0]
            self._host = {}

        def build(self, mesh, vertices_t0, vertices_t1_or_radius=None, inflation_radius=0.0, broad_phase=None):
            # build(mesh, V, r) or build(mesh, V0, V1, r) like the two C++ overloads
            counts = (C.c_int64 * 4)()
            v0, p0, ld = _f64(vertices_t0)
            if vertices_t1_or_radius is None or np.isscalar(vertices_t1_or_radius):
                r = inflation_radius if vertices_t1_or_radius is None else float(vertices_t1_or_radius)
                lib.check(lib.candidates_build_static(mesh._ctx, p0, ld, r, counts))
            else:
                v1, p1, _ = _f64(vertices_t1_or_radius)
                lib.check(lib.candidates_build_swept(mesh._ctx, p0, p1, ld, inflation_radius, counts))
            self._bind(mesh, counts)

        def _bind(self, mesh, counts):
            self.mesh = mesh
            mesh._cand_gen += 1
            self._gen = mesh._cand_gen
            self._counts = list(counts)
            self._host = {}

        def _live(self):
            if self.mesh is None or self._gen != self.mesh._cand_gen:
                raise RuntimeError("stale Candidates handle: a newer candidate set was built on this mesh")

        def _get(self, kind):
            self._live()
            if kind not in self._host:
                out = np.zeros((self._counts[kind], 2), np.int32)
                if out.size:
                    lib.check(lib.candidates_fetch(self.mesh._ctx, kind, out.ctypes.data_as(C.c_void_p)))
                self._host[kind] = out
            return self._host[kind]

        vv_candidates = property(lambda self: self._get(VV))
        ev_candidates = property(lambda self: self._get(EV))
        ee_candidates = property(lambda self: self._get(EE))
        fv_candidates = property(lambda self: self._get(FV))

        def set(self, mesh, vv=None, ev=None, ee=None, fv=None):
            """fill the container by hand (the reference's containers are public data)"""
            counts = []
            for kind, arr in enumerate((vv, ev, ee, fv)):
                a = np.ascontiguousarray(np.zeros((0, 2)) if arr is None else arr, dtype=np.int32).reshape(-1, 2)
                lib.check(lib.candidates_set(mesh._ctx, kind, a.shape[0], a.ctypes.data_as(C.c_void_p)))
                counts.append(a.shape[0])
            self._bind(mesh, counts)

        def size(self):
            return int(sum(self._counts))

        __len__ = size

        def empty(self):
            return self.size() == 0

        def compute_collision_free_stepsize(self, mesh, vertices_t0, vertices_t1, min_distance=0.0, narrow_phase_ccd=None):
            self._live()
            ccd = (narrow_phase_ccd or TightInclusionCCD())._params()
            v0, p0, ld = _f64(vertices_t0)
            v1, p1, _ = _f64(vertices_t1)
            step = C.c_double()
            lib.check(lib.ccd_stepsize_from_candidates(mesh._ctx, p0, p1, ld, min_distance, C.byref(ccd), C.byref(step)))
            return step.value

        def compute_noncandidate_conservative_stepsize(self, mesh, displacements, dhat):
            """candidates.cpp:294-338"""
            self._live()
            d, p, ld = _f64(displacements)
            step = C.c_double()
            lib.check(lib.candidates_noncandidate_stepsize(mesh._ctx, p, ld, dhat, C.byref(step)))
            return step.value

        def compute_cfl_stepsize(self, mesh, vertices_t0, vertices_t1, dhat, min_distance=0.0, broad_phase=None, narrow_phase_ccd=None):
            """candidates.cpp:340-363 (the resident candidates are replaced when the full CCD had to run)"""
            self._live()
            ccd = (narrow_phase_ccd or TightInclusionCCD())._params()
            v0, p0, ld = _f64(vertices_t0)
            v1, p1, _ = _f64(vertices_t1)
            step = C.c_double()
            lib.check(lib.candidates_cfl_stepsize(mesh._ctx, p0, p1, ld, dhat, min_distance, C.byref(ccd), C.byref(step)))
            return step.value

        def is_step_collision_free(self, mesh, vertices_t0, vertices_t1, min_distance=0.0, narrow_phase_ccd=None):
            # candidates.cpp:224-250: no candidate has an impact in [0, 1]
            return self.compute_collision_free_stepsize(mesh, vertices_t0, vertices_t1, min_distance, narrow_phase_ccd) >= 1.0

    class NormalCollisions:
        """ipc::NormalCollisions — collisions/normal/normal_collisions.cpp:20-158 (IPC and IMPROVED_MAX_APPROX set types)"""

        class CollisionSetType(enum.IntEnum):  # normal_collisions.hpp:29-39
            IPC = 0
            IMPROVED_MAX_APPROX = 1
            OGC = 2

        def __init__(self):
            self.mesh = None
            self._set = None
            self._built = False
            self._counts = [0, 0, 0, 0]
            self._host = {}
            self.use_area_weighting = False
            self.collision_set_type = NormalCollisions.CollisionSetType.IPC

        def set_use_area_weighting(self, v):
            self.use_area_weighting = bool(v)

        def set_collision_set_type(self, t):
            self.collision_set_type = NormalCollisions.CollisionSetType(t)

        def _flags(self):
            if self.collision_set_type == NormalCollisions.CollisionSetType.OGC:
                raise NotImplementedError("CollisionSetType.OGC is outside this path (DESIGN.md §7)")
            return (1 if self.use_area_weighting else 0) | (2 if self.collision_set_type else 0)

        def build(self, *args, **kw):
            """build(mesh, V, dhat, dmin=0, broad_phase=None) or build(candidates, mesh, V, dhat, dmin=0).
            defer_corrections=True (IMPROVED_MAX_APPROX built by several builders over candidate shards, include/ipcb200.h:
            IPCB_DEFER_CORRECTIONS): the build stops after this builder's sub-element pairs; exchange correction_keys() and
            finish with apply_corrections()."""
            counts = (C.c_int64 * 4)()
            flags = self._flags() | (4 if kw.get("defer_corrections") else 0)
            mesh = args[1] if isinstance(args[0], Candidates) else args[0]
            if self.mesh is not None and self.mesh is not mesh and self._set is not None:
                raise RuntimeError("a NormalCollisions object cannot move to another mesh")
            NormalCollisions._park_resident(mesh, keep=self)  # another set's records leave the context before it is overwritten
            if isinstance(args[0], Candidates):
                cand, mesh, V, dhat = args[:4]
                dmin = args[4] if len(args) > 4 else kw.get("dmin", 0.0)
                cand._live()
                v, p, ld = _f64(V)
                lib.check(lib.collisions_build_from_candidates(mesh._ctx, p, ld, dhat, dmin, flags, counts))
            else:
                mesh, V, dhat = args[:3]
                dmin = args[3] if len(args) > 3 else kw.get("dmin", 0.0)
                v, p, ld = _f64(V)
                lib.check(lib.collisions_build(mesh._ctx, p, ld, dhat, dmin, flags, counts))
                mesh._cand_gen += 1  # the resident candidates were rebuilt too
            if flags & 4:  # nothing to see yet: the records arrive with apply_corrections
                self.mesh_pending, self._pending_dmin = mesh, dmin
                return
            self._bind(mesh, counts, dmin)

        def correction_keys(self):
            """the four lists of this builder's unique sub-element pairs after a deferred build (opaque 64-bit keys)"""
            if getattr(self, "mesh_pending", None) is None:
                raise RuntimeError("no deferred build (build(..., defer_corrections=True)) on this NormalCollisions")
            n = (C.c_int64 * 4)()
            lib.check(lib.collisions_corrections_keys(self.mesh_pending._ctx, n))
            keys = np.zeros(max(1, sum(n)), np.uint64)
            if sum(n):
                lib.check(lib.collisions_corrections_pack(self.mesh_pending._ctx, keys.ctypes.data_as(C.c_void_p)))
            out, o = [], 0
            for k in range(4):
                out.append(keys[o:o + n[k]].copy())
                o += n[k]
            return out

        def apply_corrections(self, lists, rank, world):
            """lists: for each of the four lists the keys of ALL builders (concatenated, duplicates allowed); adds the
            corrections of slice `rank` of `world` and merges this builder's records"""
            mesh = getattr(self, "mesh_pending", None)
            if mesh is None:
                raise RuntimeError("no deferred build (build(..., defer_corrections=True)) on this NormalCollisions")
            n = (C.c_int64 * 4)(*[len(x) for x in lists])
            keys = np.ascontiguousarray(np.concatenate([np.asarray(x, np.uint64) for x in lists]) if sum(n) else np.zeros(1, np.uint64))
            counts = (C.c_int64 * 4)()
            lib.check(lib.collisions_corrections_apply(mesh._ctx, keys.ctypes.data_as(C.c_void_p), n, rank, world, counts))
            self._bind(mesh, counts, self._pending_dmin)
            self.mesh_pending = None

        def assign(self, mesh, builders, dmin=0.0, disjoint_shards=False):
            """Fill the set from the records of several builders and merge them like
            NormalCollisionsBuilder::merge (builder.cpp:547-689: equal collisions united, weights added,
            weight == 0 dropped).  `builders`: iterable of 4-tuples (vv, ev, ee, fv) of record namespaces as the
            *_collisions properties return them (ids, weight, eps_x, dtype) — e.g. the sets of the other ranks.
            disjoint_shards=True promises that the builders worked on disjoint candidate shards (IPCB_MERGE_DISJOINT_SHARDS)."""
            NormalCollisions._park_resident(mesh, keep=self)
            lib.check(lib.collisions_clear(mesh._ctx))
            for kinds in builders:
                for kind, rec in enumerate(kinds):
                    if rec is None:
                        continue
                    ids = np.ascontiguousarray(rec.ids, np.int32).reshape(-1, 2)
                    w = np.ascontiguousarray(rec.weight, np.float64)
                    eps = np.ascontiguousarray(rec.eps_x, np.float64)
                    dt = np.ascontiguousarray(rec.dtype, np.uint8)
                    lib.check(lib.collisions_append(mesh._ctx, kind, ids.shape[0], ids.ctypes.data_as(C.c_void_p),
                                                    w.ctypes.data_as(C.c_void_p), eps.ctypes.data_as(C.c_void_p),
                                                    dt.ctypes.data_as(C.c_void_p)))
            counts = (C.c_int64 * 4)()
            lib.check(lib.collisions_merge(mesh._ctx, dmin, 1 if disjoint_shards else 0, counts))
            self._bind(mesh, counts, dmin)

        # ---- several sets per mesh: the context works on ONE resident set; every NormalCollisions owns a collision-set
        # object of the library and is swapped in (O(1), ipcb_collisions_swap) when it is used, the previous owner's
        # records are parked in its own object first
        def _handle(self):
            if self._set is None:
                self._set = C.c_void_p()
                lib.check(lib.collision_set_create(self.mesh._ctx, C.byref(self._set)))
            return self._set

        @staticmethod
        def _park_resident(mesh, keep=None):
            owner = mesh._resident() if mesh._resident is not None else None
            if owner is not None and owner is not keep and owner.mesh is mesh:
                counts = (C.c_int64 * 4)()
                lib.check(lib.collisions_swap(mesh._ctx, owner._handle(), counts))  # the owner's records move into its object
            if owner is not keep:
                mesh._resident = None

        def _bind(self, mesh, counts, dmin):
            self.mesh = mesh
            mesh._resident = weakref.ref(self)
            self._built = True
            self._counts = list(counts)
            self._host = {}
            self.dmin = dmin

        def _live(self):
            """make this set the resident one of its mesh"""
            if self.mesh is None or not self._built:
                raise RuntimeError("NormalCollisions has not been built")
            mesh = self.mesh
            if getattr(mesh, "_ctx", None) is None:
                raise RuntimeError("the CollisionMesh of this NormalCollisions has been closed")
            if mesh._resident is not None and mesh._resident() is self:
                return
            NormalCollisions._park_resident(mesh)
            counts = (C.c_int64 * 4)()
            lib.check(lib.collisions_swap(mesh._ctx, self._handle(), counts))  # this set's records become resident
            assert list(counts) == self._counts
            mesh._resident = weakref.ref(self)

        def __del__(self):
            try:
                if self._set is not None and self.mesh is not None and getattr(self.mesh, "_ctx", None) is not None:
                    lib.collision_set_destroy(self._set)
            except Exception:  # interpreter shutdown
                pass
            self._set = None

        def _get(self, kind):
            self._live()
            if kind not in self._host:
                n = self._counts[kind]
                ids = np.zeros((n, 2), np.int32)
                w = np.zeros(n)
                eps = np.zeros(n)
                dt = np.zeros(n, np.uint8)
                if n:
                    lib.check(lib.collisions_fetch(self.mesh._ctx, kind, ids.ctypes.data_as(C.c_void_p),
                                                   w.ctypes.data_as(C.c_void_p), eps.ctypes.data_as(C.c_void_p),
                                                   dt.ctypes.data_as(C.c_void_p)))
                self._host[kind] = types.SimpleNamespace(ids=ids, weight=w, eps_x=eps, dtype=dt)
            return self._host[kind]

        vv_collisions = property(lambda self: self._get(VV))
        ev_collisions = property(lambda self: self._get(EV))
        ee_collisions = property(lambda self: self._get(EE))
        fv_collisions = property(lambda self: self._get(FV))

        def counts(self):
            return list(self._counts)

        def size(self):
            return int(sum(self._counts))

        __len__ = size

        def empty(self):
            return self.size() == 0

        def compute_minimum_distance(self, mesh, vertices):
            self._live()
            v, p, ld = _f64(vertices)
            out = C.c_double()
            lib.check(lib.collisions_min_distance(mesh._ctx, p, ld, C.byref(out)))
            return out.value

    class BarrierPotential:
        """ipc::BarrierPotential(dhat, stiffness, use_physical_barrier) — potentials/potential.cpp:36-222"""

        def __init__(self, dhat, stiffness=1.0, use_physical_barrier=False):
            self.dhat = dhat
            self.stiffness = stiffness
            self.use_physical_barrier = use_physical_barrier

        def _bp(self):
            return _abi.BarrierParams(self.dhat, self.stiffness, int(self.use_physical_barrier))

        def __call__(self, collisions, mesh, X):
            collisions._live()
            x, p, ld = _f64(X)
            e = C.c_double()
            bp = self._bp()
            lib.check(lib.barrier_energy(mesh._ctx, p, ld, C.byref(bp), C.byref(e)))
            return e.value

        def gradient(self, collisions, mesh, X):
            collisions._live()
            x, p, ld = _f64(X)
            g = np.zeros(3 * mesh.num_vertices())
            bp = self._bp()
            lib.check(lib.barrier_gradient(mesh._ctx, p, ld, C.byref(bp), g.ctypes.data_as(C.c_void_p)))
            return g

        def hessian(self, collisions, mesh, X, project_hessian_to_psd=PSDProjectionMethod.NONE):
            import scipy.sparse as sp

            collisions._live()
            x, p, ld = _f64(X)
            nnz = C.c_int64()
            bp = self._bp()
            lib.check(lib.barrier_hessian(mesh._ctx, p, ld, C.byref(bp), int(project_hessian_to_psd), C.byref(nnz)))
            n = 3 * mesh.num_vertices()
            outer = np.zeros(n + 1, np.int32)
            inner = np.zeros(nnz.value, np.int32)
            vals = np.zeros(nnz.value)
            lib.check(lib.barrier_hessian_fetch(mesh._ctx, outer.ctypes.data_as(C.c_void_p),
                                                inner.ctypes.data_as(C.c_void_p), vals.ctypes.data_as(C.c_void_p)))
            return sp.csc_matrix((vals, inner, outer), shape=(n, n))

    class TangentialCollisions:
        """ipc::TangentialCollisions — collisions/tangential/tangential_collisions.cpp:62-171 (isotropic coefficients).
        One resident tangential set per mesh (the lagged set of a friction solve)."""

        def __init__(self):
            self.mesh = None
            self._counts = [0, 0, 0, 0]
            self._host = {}

        def build(self, mesh, vertices, collisions, normal_potential, mu_s, mu_k=None):
            """mu_s / mu_k: a scalar or one value per vertex (mu_k defaults to mu_s, like the single-mu overload)"""
            collisions._live()
            nV = mesh.num_vertices()
            ms = np.ascontiguousarray(np.broadcast_to(np.asarray(mu_s, np.float64), (nV,)))
            mk = ms if mu_k is None else np.ascontiguousarray(np.broadcast_to(np.asarray(mu_k, np.float64), (nV,)))
            v, p, ld = _f64(vertices)
            bp = normal_potential._bp()
            counts = (C.c_int64 * 4)()
            lib.check(lib.tangential_build(mesh._ctx, p, ld, C.byref(bp), ms.ctypes.data_as(C.c_void_p), mk.ctypes.data_as(C.c_void_p), counts))
            self.mesh, self._counts, self._host = mesh, list(counts), {}
            mesh._tang_gen = getattr(mesh, "_tang_gen", 0) + 1
            self._gen = mesh._tang_gen

        def _live(self):
            if self.mesh is None or self._gen != self.mesh._tang_gen:
                raise RuntimeError("stale TangentialCollisions handle: a newer tangential set was built on this mesh")

        def _get(self, kind):
            self._live()
            if kind not in self._host:
                n = self._counts[kind]
                r = types.SimpleNamespace(ids=np.zeros((n, 2), np.int32), weight=np.zeros(n), normal_force_magnitude=np.zeros(n), mu_s=np.zeros(n),
                                          mu_k=np.zeros(n), closest_point=np.zeros((n, 2)), tangent_basis=np.zeros((n, 2, 3)))
                if n:
                    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
                    lib.check(lib.tangential_fetch(self.mesh._ctx, kind, ptr(r.ids), ptr(r.weight), ptr(r.normal_force_magnitude), ptr(r.mu_s),
                                                   ptr(r.mu_k), ptr(r.closest_point), ptr(r.tangent_basis)))
                self._host[kind] = r
            return self._host[kind]

        vv_collisions = property(lambda self: self._get(VV))
        ev_collisions = property(lambda self: self._get(EV))
        ee_collisions = property(lambda self: self._get(EE))
        fv_collisions = property(lambda self: self._get(FV))

        def counts(self):
            return list(self._counts)

        def size(self):
            return int(sum(self._counts))

        __len__ = size

        def empty(self):
            return self.size() == 0

    class FrictionPotential:
        """ipc::FrictionPotential(eps_v) — potentials/friction_potential.hpp, tangential_potential.cpp:162-325"""

        def __init__(self, eps_v):
            if not eps_v > 0:
                raise ValueError("eps_v must be positive")
            self.eps_v = float(eps_v)

        def __call__(self, collisions, mesh, velocities):
            collisions._live()
            x, p, ld = _f64(velocities)
            e = C.c_double()
            lib.check(lib.friction_energy(mesh._ctx, p, ld, self.eps_v, C.byref(e)))
            return e.value

        def gradient(self, collisions, mesh, velocities):
            collisions._live()
            x, p, ld = _f64(velocities)
            g = np.zeros(3 * mesh.num_vertices())
            lib.check(lib.friction_gradient(mesh._ctx, p, ld, self.eps_v, g.ctypes.data_as(C.c_void_p)))
            return g

        def hessian(self, collisions, mesh, velocities, project_hessian_to_psd=PSDProjectionMethod.NONE):
            import scipy.sparse as sp

            collisions._live()
            x, p, ld = _f64(velocities)
            nnz = C.c_int64()
            lib.check(lib.friction_hessian(mesh._ctx, p, ld, self.eps_v, int(project_hessian_to_psd), C.byref(nnz)))
            n = 3 * mesh.num_vertices()
            outer, inner, vals = np.zeros(n + 1, np.int32), np.zeros(nnz.value, np.int32), np.zeros(nnz.value)
            lib.check(lib.barrier_hessian_fetch(mesh._ctx, outer.ctypes.data_as(C.c_void_p), inner.ctypes.data_as(C.c_void_p),
                                                vals.ctypes.data_as(C.c_void_p)))
            return sp.csc_matrix((vals, inner, outer), shape=(n, n))

    def compute_collision_free_stepsize(mesh, vertices_t0, vertices_t1, min_distance=0.0, broad_phase=None,
                                        narrow_phase_ccd=None):
        """ipc::compute_collision_free_stepsize — ipc.cpp:45-101"""
        ccd = (narrow_phase_ccd or TightInclusionCCD())._params()
        v0, p0, ld = _f64(vertices_t0)
        v1, p1, _ = _f64(vertices_t1)
        step = C.c_double()
        lib.check(lib.ccd_stepsize(mesh._ctx, p0, p1, ld, min_distance, C.byref(ccd), C.byref(step)))
        mesh._cand_gen += 1
        return step.value

    def is_step_collision_free(mesh, vertices_t0, vertices_t1, min_distance=0.0, broad_phase=None, narrow_phase_ccd=None):
        """ipc::is_step_collision_free — ipc.cpp:20-43"""
        return compute_collision_free_stepsize(mesh, vertices_t0, vertices_t1, min_distance, broad_phase,
                                               narrow_phase_ccd) >= 1.0

    def has_intersections(mesh, vertices, broad_phase=None):
        """ipc::has_intersections — ipc.cpp:105-166 (3D: an edge crossing a triangle)"""
        v, p, ld = _f64(vertices)
        out = C.c_int32()
        lib.check(lib.has_intersections(mesh._ctx, p, ld, C.byref(out)))
        mesh._cand_gen += 1
        return bool(out.value)

    def narrow_phase_ccd(kind, x_t0, x_t1, min_distance=0.0, tmax=1.0, ccd=None, mesh=None):
        """batched NarrowPhaseCCD queries (ccd/narrow_phase_ccd.hpp:8-119): x_* are (n, 4, 3) arrays"""
        a = np.ascontiguousarray(x_t0, dtype=np.float64).reshape(-1, 12)
        b = np.ascontiguousarray(x_t1, dtype=np.float64).reshape(-1, 12)
        n = a.shape[0]
        hit = np.zeros(n, np.uint8)
        toi = np.zeros(n)
        params = (ccd or TightInclusionCCD())._params()
        own = None
        if mesh is None:
            own = CollisionMesh(np.zeros((1, 3)))
            mesh = own
        lib.check(lib.ccd_narrow_phase(mesh._ctx, kind, n, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p),
                                       min_distance, tmax, C.byref(params), hit.ctypes.data_as(C.c_void_p),
                                       toi.ctypes.data_as(C.c_void_p)))
        return hit.astype(bool), toi

    ns.CollisionMesh = CollisionMesh
    ns.BroadPhase = BroadPhase
    ns.Candidates = Candidates
    ns.NormalCollisions = NormalCollisions
    ns.BarrierPotential = BarrierPotential
    ns.TangentialCollisions = TangentialCollisions
    ns.FrictionPotential = FrictionPotential
    ns.compute_collision_free_stepsize = compute_collision_free_stepsize
    ns.is_step_collision_free = is_step_collision_free
    ns.narrow_phase_ccd = narrow_phase_ccd
    ns.has_intersections = has_intersections
    return ns
