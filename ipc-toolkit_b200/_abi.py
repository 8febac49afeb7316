"""ctypes declarations for the C ABI in include/ipcb200.h.

The declarations are keyed by a symbol prefix: the product library uses
``ipcb_``; the test-only CPU checker binds the host half of the same table
with its own prefix from its own directory.  Nothing here computes anything.
"""
import ctypes as C

c_i32, c_i64, c_f64 = C.c_int32, C.c_int64, C.c_double
P = C.c_void_p


class CcdParams(C.Structure):
    _fields_ = [("kind", c_i32), ("tolerance", c_f64), ("max_iterations", c_i64), ("conservative_rescaling", c_f64)]


class BarrierParams(C.Structure):
    _fields_ = [("dhat", c_f64), ("stiffness", c_f64), ("use_physical_barrier", c_i32)]


# name -> (restype, argtypes); ctx is always the first void*
HOST_API = {
    "ctx_create": (C.c_int, [C.c_int, C.POINTER(P)]),
    "ctx_destroy": (None, [P]),
    "last_error": (C.c_char_p, []),
    "backend_name": (C.c_char_p, []),
    "ctx_stream": (P, [P]),
    "mesh_set": (C.c_int, [P, c_i32, P, c_i32, c_i32, P, c_i32, c_i32, P, c_i32]),
    "mesh_set_collision_filter": (C.c_int, [P, P, c_i32]),
    "mesh_num_codim_vertices": (C.c_int, [P, C.POINTER(c_i32)]),
    "mesh_num_codim_edges": (C.c_int, [P, C.POINTER(c_i32)]),
    "mesh_faces_to_edges": (C.c_int, [P, P]),
    "mesh_areas": (C.c_int, [P, P, P]),
    "broad_build_static": (C.c_int, [P, P, c_i32, c_f64, c_i32]),
    "broad_build_swept": (C.c_int, [P, P, P, c_i32, c_f64, c_i32]),
    "broad_detect": (C.c_int, [P, c_i32, C.POINTER(c_i64)]),
    "broad_fetch": (C.c_int, [P, c_i32, P]),
    "broad_vertex_boxes": (C.c_int, [P, P]),
    "candidates_build_static": (C.c_int, [P, P, c_i32, c_f64, C.POINTER(c_i64)]),
    "candidates_build_swept": (C.c_int, [P, P, P, c_i32, c_f64, C.POINTER(c_i64)]),
    "candidates_fetch": (C.c_int, [P, c_i32, P]),
    "candidates_set": (C.c_int, [P, c_i32, c_i64, P]),
    "collisions_build": (C.c_int, [P, P, c_i32, c_f64, c_f64, c_i32, C.POINTER(c_i64)]),
    "collisions_build_from_candidates": (C.c_int, [P, P, c_i32, c_f64, c_f64, c_i32, C.POINTER(c_i64)]),
    "collisions_fetch": (C.c_int, [P, c_i32, P, P, P, P]),
    "collisions_min_distance": (C.c_int, [P, P, c_i32, C.POINTER(c_f64)]),
    "collisions_clear": (C.c_int, [P]),
    "collisions_append": (C.c_int, [P, c_i32, c_i64, P, P, P, P]),
    "collisions_merge": (C.c_int, [P, c_f64, c_i32, C.POINTER(c_i64)]),
    "collision_set_create": (C.c_int, [P, C.POINTER(P)]),
    "collision_set_destroy": (None, [P]),
    "collisions_swap": (C.c_int, [P, P, C.POINTER(c_i64)]),
    "ctx_set_collision_range": (C.c_int, [P, c_i32, c_i32]),
    "ctx_set_row_block": (C.c_int, [P, c_i32, c_i32]),
    "hessian_balanced_row_blocks": (C.c_int, [P, c_i32, P]),
    "collisions_corrections_keys": (C.c_int, [P, C.POINTER(c_i64)]),
    "collisions_corrections_pack": (C.c_int, [P, P]),
    "collisions_corrections_apply": (C.c_int, [P, P, C.POINTER(c_i64), c_i32, c_i32, C.POINTER(c_i64)]),
    "barrier_energy": (C.c_int, [P, P, c_i32, C.POINTER(BarrierParams), C.POINTER(c_f64)]),
    "barrier_gradient": (C.c_int, [P, P, c_i32, C.POINTER(BarrierParams), P]),
    "barrier_hessian": (C.c_int, [P, P, c_i32, C.POINTER(BarrierParams), c_i32, C.POINTER(c_i64)]),
    "barrier_hessian_fetch": (C.c_int, [P, P, P, P]),
    "has_intersections": (C.c_int, [P, P, c_i32, C.POINTER(c_i32)]),
    "tangential_build": (C.c_int, [P, P, c_i32, C.POINTER(BarrierParams), P, P, C.POINTER(c_i64)]),
    "tangential_fetch": (C.c_int, [P, c_i32, P, P, P, P, P, P, P]),
    "friction_energy": (C.c_int, [P, P, c_i32, c_f64, C.POINTER(c_f64)]),
    "friction_gradient": (C.c_int, [P, P, c_i32, c_f64, P]),
    "friction_hessian": (C.c_int, [P, P, c_i32, c_f64, c_i32, C.POINTER(c_i64)]),
    "ccd_stepsize": (C.c_int, [P, P, P, c_i32, c_f64, C.POINTER(CcdParams), C.POINTER(c_f64)]),
    "ccd_stepsize_from_candidates": (C.c_int, [P, P, P, c_i32, c_f64, C.POINTER(CcdParams), C.POINTER(c_f64)]),
    "candidates_noncandidate_stepsize": (C.c_int, [P, P, c_i32, c_f64, C.POINTER(c_f64)]),
    "candidates_cfl_stepsize": (C.c_int, [P, P, P, c_i32, c_f64, c_f64, C.POINTER(CcdParams), C.POINTER(c_f64)]),
    "ccd_narrow_phase": (C.c_int, [P, c_i32, c_i64, P, P, c_f64, c_f64, C.POINTER(CcdParams), P, P]),
}

# device-resident variants: product only
DEVICE_API = {
    "tangential_build_dev": (C.c_int, [P, P, c_i32, C.POINTER(BarrierParams), P, P, C.POINTER(c_i64)]),
    "friction_energy_dev": (C.c_int, [P, P, c_i32, c_f64, P]),
    "friction_gradient_dev": (C.c_int, [P, P, c_i32, c_f64, P]),
    "friction_hessian_dev": (C.c_int, [P, P, c_i32, c_f64, c_i32, C.POINTER(c_i64)]),
    "collisions_build_dev": (C.c_int, [P, P, c_i32, c_f64, c_f64, c_i32, C.POINTER(c_i64)]),
    "collisions_build_from_candidates_dev": (C.c_int, [P, P, c_i32, c_f64, c_f64, c_i32, C.POINTER(c_i64)]),
    "barrier_energy_dev": (C.c_int, [P, P, c_i32, C.POINTER(BarrierParams), P]),
    "barrier_gradient_dev": (C.c_int, [P, P, c_i32, C.POINTER(BarrierParams), P]),
    "barrier_hessian_dev": (C.c_int, [P, P, c_i32, C.POINTER(BarrierParams), c_i32, C.POINTER(c_i64)]),
    "barrier_hessian_dev_ptrs": (C.c_int, [P, C.POINTER(P), C.POINTER(P), C.POINTER(P)]),
    "collisions_dev_ptrs": (C.c_int, [P, c_i32, C.POINTER(c_i64), C.POINTER(P), C.POINTER(P), C.POINTER(P), C.POINTER(P)]),
    "collisions_append_dev": (C.c_int, [P, c_i32, c_i64, P, P, P, P]),
    "collisions_corrections_keys_dev": (C.c_int, [P, C.POINTER(c_i64)]),
    "collisions_corrections_pack_dev": (C.c_int, [P, P]),
    "collisions_corrections_apply_dev": (C.c_int, [P, P, C.POINTER(c_i64), C.POINTER(c_i64)]),
    "collisions_pack_dev": (C.c_int, [P, P, c_i64, C.POINTER(c_i64)]),
    "collisions_append_packed_dev": (C.c_int, [P, P, C.POINTER(c_i64)]),
    "candidates_build_swept_dev": (C.c_int, [P, P, P, c_i32, c_f64, C.POINTER(c_i64)]),
    "ccd_stepsize_dev": (C.c_int, [P, P, P, c_i32, c_f64, C.POINTER(CcdParams), P]),
    "ccd_stepsize_from_candidates_dev": (C.c_int, [P, P, P, c_i32, c_f64, C.POINTER(CcdParams), P]),
    "ctx_set_shard": (C.c_int, [P, c_i32, c_i32]),
    "ctx_set_broad_phase_method": (C.c_int, [P, c_i32]),
    "ctx_launch_count": (C.c_int, [P, C.POINTER(c_i64)]),
    "ctx_enable_stage_timing": (C.c_int, [P, c_i32]),
    "ctx_stage_times": (C.c_int, [P, c_i32, C.POINTER(C.c_char_p), C.POINTER(C.c_float)]),
    "measure_fp64_peak": (C.c_int, [P, c_i32, C.POINTER(c_f64)]),
    "measure_copy_bandwidth": (C.c_int, [P, c_i64, c_i32, C.POINTER(c_f64)]),
}


class Lib:
    """A loaded library exposing ``self.<name>`` for every ABI function."""

    def __init__(self, path, prefix, device_api):
        self.path = str(path)
        self.prefix = prefix
        self.cdll = C.CDLL(self.path)
        self.has_device_api = device_api
        table = dict(HOST_API)
        if device_api:
            table.update(DEVICE_API)
        self.names = sorted(table)
        for name, (res, args) in table.items():
            fn = getattr(self.cdll, prefix + name)  # AttributeError if a declared symbol is missing
            fn.restype = res
            fn.argtypes = args
            setattr(self, name, fn)

    def check(self, rc):
        if rc != 0:
            raise RuntimeError(self.last_error().decode() or "unknown error in %s" % self.path)

    def backend(self):
        return self.backend_name().decode()
