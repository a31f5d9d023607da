"""ctypes binding of include/miniaero_b200.h (the C ABI).  No torch types cross this boundary.

The shared library is built in-tree by `miniaero_b200/build.py` (called from `__graft_entry__.build()`).
There is no fallback: if the library is missing, importing fails loudly.
"""
import ctypes as C
import os

MA_MAX_BC_SETS = 16
MA_COMM_ID_BYTES = 128

MA_OK = 0
BC_EXTRAPOLATE, BC_TANGENT, BC_INFLOW, BC_NOSLIP = 0, 1, 2, 3
BC_NAMES = {"Extrapolate": BC_EXTRAPOLATE, "Tangent": BC_TANGENT, "Inflow": BC_INFLOW, "NoSlip": BC_NOSLIP}
ARITH_FAST, ARITH_STRICT = 0, 1
FIELD_GRADIENT, FIELD_LIMITER, FIELD_STAGE_PRIMITIVES = 0, 1, 2
LIMITER_VENKAT, LIMITER_VANALBADA = 0, 1

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class Options(C.Structure):
    """ma_options == the reference's `struct Options` (Options.h:47-57)."""
    _fields_ = [("problem_type", C.c_int), ("lx", C.c_double), ("ly", C.c_double), ("lz", C.c_double),
                ("angle", C.c_double), ("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int),
                ("ntimesteps", C.c_int), ("dt", C.c_double), ("output_results", C.c_int),
                ("output_frequency", C.c_int), ("second_order_space", C.c_int), ("viscous", C.c_int)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Faces(C.Structure):
    _fields_ = [("nfaces", C.c_int), ("coordinates", _dp), ("face_normal", _dp), ("face_tangent", _dp),
                ("face_binormal", _dp), ("face_cell_conn", _ip), ("cell_flux_index", _ip)]


class Mesh(C.Structure):
    _fields_ = [("num_owned_cells", C.c_int), ("num_ghosts", C.c_int), ("cell_coordinates", _dp),
                ("cell_volumes", _dp), ("internal_faces", Faces), ("num_boundary_sets", C.c_int),
                ("boundary_type", C.c_int * MA_MAX_BC_SETS), ("boundary_faces", Faces * MA_MAX_BC_SETS),
                ("num_ranks", C.c_int), ("my_rank", C.c_int), ("send_count", _ip), ("recv_count", _ip),
                ("send_local_ids", _ip), ("recv_local_ids", _ip)]


class SolverConfig(C.Structure):
    _fields_ = [("device", C.c_int), ("arith", C.c_int), ("tile_dims", C.c_int * 3), ("block_threads", C.c_int),
                ("comm", C.c_void_p), ("overlap_halo", C.c_int), ("stream", C.c_void_p), ("limiter", C.c_int),
                ("share_cut_faces", C.c_int)]


class Timing(C.Structure):
    _fields_ = [("step_seconds", C.c_double), ("steps", C.c_longlong), ("cell_updates", C.c_longlong),
                ("grad_seconds", C.c_double), ("flux_seconds", C.c_double), ("halo_seconds", C.c_double),
                ("kernel_launches", C.c_longlong), ("device_bytes", C.c_size_t), ("num_tiles", C.c_int),
                ("tile_faces_total", C.c_int), ("num_interior_tiles", C.c_int), ("num_send_cells", C.c_int),
                ("num_recv_cells", C.c_int), ("halo_wait_seconds", C.c_double), ("faces_evaluated", C.c_longlong)]


class Report(C.Structure):
    """ma_report: the fields of the Mantevo YAML run report (host driver)."""
    _fields_ = [("app_name", C.c_char_p), ("app_version", C.c_char_p), ("options", C.POINTER(Options)),
                ("num_ranks", C.c_int), ("blocks", C.c_int * 3), ("global_cells", C.c_longlong),
                ("timing", C.POINTER(Timing)), ("setup_seconds", C.c_double), ("run_seconds", C.c_double),
                ("total_seconds", C.c_double), ("hbm_peak_gbs", C.c_double), ("device_name", C.c_char_p)]


LIB_NAME = "libminiaero_b200.so"
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), LIB_NAME)
# developer knob (tools/build_variants.py): load an experiment build of the same library instead
if os.environ.get("MINIAERO_B200_LIB"):
    LIB_PATH = os.path.abspath(os.environ["MINIAERO_B200_LIB"])

# every symbol include/miniaero_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "ma_last_error": (C.c_char_p, []),
    "ma_abi_version": (C.c_int, []),
    "ma_solver_debug_array": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "ma_options_default": (None, [C.POINTER(Options)]),
    "ma_options_read": (C.c_int, [C.c_char_p, C.POINTER(Options)]),
    "ma_mesh_generate": (C.c_int, [C.POINTER(Options), C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "ma_mesh_view": (C.POINTER(Mesh), [C.c_void_p]),
    "ma_mesh_global_ids": (_ip, [C.c_void_p]),
    "ma_mesh_decomposition": (None, [C.c_void_p, _ip, _ip, _ip, _ip]),
    "ma_mesh_free": (None, [C.c_void_p]),
    "ma_block_decomposition": (C.c_int, [C.POINTER(Options), C.c_int, C.c_int, _ip, _ip, _ip, _ip]),
    "ma_comm_get_unique_id": (C.c_int, [C.c_char_p]),
    "ma_comm_create": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "ma_comm_destroy": (None, [C.c_void_p]),
    "ma_solver_config_default": (None, [C.POINTER(SolverConfig)]),
    "ma_solver_create": (C.c_int, [C.POINTER(Mesh), C.POINTER(Options), C.POINTER(SolverConfig),
                                   C.POINTER(C.c_void_p)]),
    "ma_solver_create_structured": (C.c_int, [C.POINTER(Options), C.c_int, C.c_int, C.POINTER(SolverConfig),
                                              C.POINTER(C.c_void_p)]),
    "ma_solver_num_cells": (C.c_int, [C.c_void_p, _ip, _ip]),
    "ma_solver_destroy": (None, [C.c_void_p]),
    "ma_solver_initialize": (C.c_int, [C.c_void_p]),
    "ma_solver_step": (C.c_int, [C.c_void_p, C.c_int]),
    "ma_solver_solve": (C.c_int, [C.c_void_p]),
    "ma_solver_synchronize": (C.c_int, [C.c_void_p]),
    "ma_solver_get_solution": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ma_solver_set_solution": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ma_solver_submit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "ma_solver_get_field": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "ma_solver_get_timing": (C.c_int, [C.c_void_p, C.POINTER(Timing)]),
    "ma_solver_reset_timing": (C.c_int, [C.c_void_p]),
    "ma_solver_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "ma_write_results": (C.c_int, [C.c_char_p, C.POINTER(Mesh), C.c_void_p, C.c_int]),
    "ma_write_yaml_report": (C.c_int, [C.POINTER(Report), C.c_char_p, C.c_char_p, C.c_size_t]),
    "ma_probe_roe_flux": (C.c_int, [C.c_int] + [C.c_void_p] * 6 + [C.c_int, C.c_int]),
    "ma_probe_viscous_flux": (C.c_int, [C.c_int] + [C.c_void_p] * 4 + [C.c_int, C.c_int]),
    "ma_probe_primitives": (C.c_int, [C.c_int] + [C.c_void_p] * 2 + [C.c_int, C.c_int]),
    "ma_probe_venkat": (C.c_int, [C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.c_int]),
    "ma_probe_vanalbada": (C.c_int, [C.c_int] + [C.c_void_p] * 4 + [C.c_int, C.c_int]),
}

_lib = None


class MiniAeroError(RuntimeError):
    pass


def load():
    """dlopen the in-tree C-ABI library and type every entry point.  Fails loudly when absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise MiniAeroError("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code):
    if code != MA_OK:
        msg = load().ma_last_error()
        raise MiniAeroError("miniaero_b200 error %d: %s" % (code, msg.decode() if msg else "?"))
