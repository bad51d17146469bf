"""ctypes binding of libb200stencil.so -- the same C ABI (include/b200stencil.h) the Julia shims bind with ccall.

There is NO CPU fallback: if the shared library is missing this module raises at import of the first symbol, and
every compute entry point returns an error status when no CUDA device is present.
"""
import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200stencil.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "b200stencil.h")

# ---- constants (mirror include/b200stencil.h) ---------------------------------------------------------------------
OK, ERR_BAD_SIZE, ERR_BAD_ARG, ERR_CUDA, ERR_NOT_IMPLEMENTED, ERR_NO_DEVICE, ERR_STATE = range(7)
HALO_REFERENCE_LAG2, HALO_CONSISTENT = 0, 1
BC_LITERAL, BC_PROPER = 0, 1
KERNEL_AUTO, KERNEL_DIRECT, KERNEL_TMA = 0, 1, 2
ARITH_KERNEL, ARITH_ARRAY = 0, 1
COARSE_JACOBI, COARSE_CG = 0, 1
SMOOTH_JACOBI, SMOOTH_RBGS = 0, 1
RESTRICT_INJECT, RESTRICT_FW = 0, 1
POLICY_SERIAL, POLICY_PARALLEL, POLICY_PARALLEL_SHMEM = 0, 1, 2
PCG_TOL_INITIAL_RESIDUAL, PCG_TOL_RHS = 0, 1
NS_SOLVER_VCYCLE, NS_SOLVER_MG_PCG = 0, 1

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_vp = C.c_void_p


class B2SError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"b200stencil error {code}: {msg}")
        self.code = code


class Diff3DConfig(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("nslabs_total", C.c_int), ("slab_begin", C.c_int),
                ("slab_count", C.c_int), ("devices", _ip), ("halo_mode", C.c_int), ("bc_mode", C.c_int),
                ("scale_physical_size", C.c_int), ("kernel_variant", C.c_int), ("batch", C.c_int),
                ("dimx", C.c_int), ("dimy", C.c_int), ("arithmetic", C.c_int)]


class Diff3DParams(C.Structure):
    _fields_ = [("lx", C.c_double), ("ly", C.c_double), ("lz", C.c_double), ("dx", C.c_double), ("dy", C.c_double),
                ("dz", C.c_double), ("dt", C.c_double), ("dtau", C.c_double), ("total_N", C.c_double),
                ("nx_g", C.c_int), ("ny_g", C.c_int), ("nz_g", C.c_int)]


class MGConfig(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("coarse_solve_size", C.c_int), ("coarse_solver", C.c_int),
                ("smoother", C.c_int), ("restriction", C.c_int), ("device", C.c_int), ("use_graph", C.c_int),
                ("smem_levels", C.c_int), ("fuse_sweeps", C.c_int)]


class NS2DParams(C.Structure):
    _fields_ = [("k", C.c_double), ("Ra", C.c_double), ("Pr", C.c_double), ("nx", C.c_int), ("ny", C.c_int),
                ("ttot", C.c_double), ("beta", C.c_double), ("niters", C.c_int), ("tol", C.c_double),
                ("a_dif", C.c_double), ("a_adv", C.c_double)]


class NS2DStepInfo(C.Structure):
    _fields_ = [("dt", C.c_double), ("cycles_S", C.c_int), ("cycles_T", C.c_int), ("cycles_W", C.c_int),
                ("r_S", C.c_double), ("r_T", C.c_double), ("r_W", C.c_double)]


_d, _i, _sz, _ll = C.c_double, C.c_int, C.c_size_t, C.c_longlong
_llp = C.POINTER(C.c_longlong)

# name -> (restype, argtypes). Every function declared in include/b200stencil.h must appear here (tests check it).
SIGNATURES = {
    "b2s_last_error": (C.c_char_p, []),
    "b2s_version": (_i, []),
    "b2s_device_count": (_i, [_ip]),
    "b2s_shutdown": (_i, []),
    "b2s_diffusion3d_step_tau": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _d, _d, _d, _d, _d, _d, _d, _d, _d, _vp, _i, _vp]),
    "b2s_diff3d_create": (_i, [C.POINTER(_vp), C.POINTER(Diff3DConfig)]),
    "b2s_diff3d_destroy": (_i, [_vp]),
    "b2s_diff3d_get_params": (_i, [_vp, C.POINTER(Diff3DParams)]),
    "b2s_diff3d_params_for": (_i, [C.POINTER(Diff3DConfig), C.POINTER(Diff3DParams)]),
    "b2s_diff3d_init_gaussian": (_i, [_vp]),
    "b2s_diff3d_set_initial": (_i, [_vp, _vp]),
    "b2s_diff3d_ipc_blob_bytes": (_sz, []),
    "b2s_diff3d_ipc_export": (_i, [_vp, _vp]),
    "b2s_diff3d_ipc_connect": (_i, [_vp, _vp, _i]),
    "b2s_diff3d_solve_timestep": (_i, [_vp, _d, _i, _ip, _dp]),
    "b2s_diff3d_iterate": (_i, [_vp, _i, _vp]),
    "b2s_diff3d_advance_time": (_i, [_vp]),
    "b2s_diff3d_run": (_i, [_vp, _d, _d, _i, _ip, _i, _ip]),
    "b2s_diff3d_get_field": (_i, [_vp, _i, _i, _vp]),
    "b2s_diff3d_gather": (_i, [_vp, _vp]),
    "b2s_diff3d_device_ptr": (_i, [_vp, _i, _i, C.POINTER(_vp)]),
    "b2s_diff3d_upload_state": (_i, [_vp, _i, _vp]),
    "b2s_diff3d_download_state": (_i, [_vp, _i, _vp]),
    "b2s_diff3d_download_state_async": (_i, [_vp, _i, _vp]),
    "b2s_diff3d_upload_state_async": (_i, [_vp, _i, _vp]),
    "b2s_diff3d_commit_upload": (_i, [_vp, _i]),
    "b2s_diff3d_sync": (_i, [_vp]),
    "b2s_diff3d_stats": (_i, [_vp, _llp, _dp]),
    "b2s_residual2d": (_i, [_vp, _vp, _d, _d, _vp, _i, _i, _i, _vp]),
    "b2s_iteration2d": (_i, [_vp, _vp, _d, _d, _vp, _i, _i, _d, _i, _dp, _vp]),
    "b2s_rbgs2d": (_i, [_vp, _vp, _d, _d, _i, _i, _dp, _vp]),
    "b2s_restrict_inject2d": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "b2s_restrict_fw2d": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "b2s_prolongate2d": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "b2s_matvec2d": (_i, [_vp, _d, _d, _d, _vp, _i, _i, _i, _vp]),
    "b2s_apply_bc2d": (_i, [_vp, _i, _i, _i, _vp]),
    "b2s_dot": (_i, [_vp, _vp, _sz, _dp, _vp]),
    "b2s_sumsq": (_i, [_vp, _sz, _dp, _vp]),
    "b2s_axpy": (_i, [_d, _vp, _vp, _sz, _vp]),
    "b2s_xpby": (_i, [_vp, _d, _vp, _sz, _vp]),
    "b2s_mg_create": (_i, [C.POINTER(_vp), C.POINTER(MGConfig)]),
    "b2s_mg_destroy": (_i, [_vp]),
    "b2s_mg_solve": (_i, [_vp, _vp, _vp, _d, _d, _d, _i, _i, _dp, _ip, _vp]),
    "b2s_mg_vcycle": (_i, [_vp, _vp, _vp, _d, _d, _d, _i, _dp]),
    "b2s_mg_cycles": (_i, [_vp, _vp, _vp, _d, _d, _d, _i, _i, _dp, _dp]),
    "b2s_mg_pcg_solve": (_i, [_vp, _vp, _vp, _d, _d, _d, _i, _dp, _ip]),
    "b2s_mg_pcg_solve2": (_i, [_vp, _vp, _vp, _d, _d, _d, _i, _i, _dp, _ip]),
    "b2s_mg_profile_kernels": (_i, [_vp, _vp, _vp, _d, _d, _i, _ip, _dp, _dp, _dp, _ip, _ip]),
    "b2s_mg_last_coarse_sweeps": (_i, [_vp, _ip]),
    "b2s_mg_stats": (_i, [_vp, _llp, _dp]),
    "b2s_cg_solve": (_i, [_vp, _vp, _d, _d, _d, _d, _i, _i, _i, _i, _dp, _ip, _vp]),
    "b2s_ns2d_create": (_i, [C.POINTER(_vp), C.POINTER(NS2DParams), C.POINTER(MGConfig)]),
    "b2s_ns2d_destroy": (_i, [_vp]),
    "b2s_ns2d_set_solver": (_i, [_vp, _i]),
    "b2s_ns2d_set_field": (_i, [_vp, _i, _vp]),
    "b2s_ns2d_get_field": (_i, [_vp, _i, _vp]),
    "b2s_ns2d_init_cosine": (_i, [_vp, _i]),
    "b2s_ns2d_step": (_i, [_vp, C.POINTER(NS2DStepInfo)]),
    "b2s_ns2d_get_aux": (_i, [_vp, _i, _vp]),
}


def header_symbols(path=HEADER_PATH):
    """Function names declared in include/b200stencil.h."""
    txt = open(path).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b2s_[a-z0-9_]+)\s*\(", txt)))


_LIB = None


def lib():
    """Loads libb200stencil.so (fails loudly when it has not been built: run __graft_entry__.build())."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                              "g.build()'` (make -C finalprojectrepo.jl_b200/csrc). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(L, name)
            except AttributeError:  # reported by missing_symbols(); calling it raises AttributeError
                continue
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def missing_symbols():
    """Header-declared functions the built library does not export (must be empty)."""
    L = lib()
    return [n for n in header_symbols() if not hasattr(L, n)]


def check(code):
    if code != OK:
        raise B2SError(code, lib().b2s_last_error().decode(errors="replace"))


def device_count():
    n = C.c_int(0)
    rc = lib().b2s_device_count(C.byref(n))
    return n.value if rc == OK else 0


def ptr(x):
    """Raw address of a torch CUDA tensor / numpy array / int / None as c_void_p."""
    if x is None:
        return None
    if isinstance(x, int):
        return C.c_void_p(x)
    if hasattr(x, "data_ptr"):
        return C.c_void_p(x.data_ptr())
    if hasattr(x, "ctypes"):
        return C.c_void_p(x.ctypes.data)
    raise TypeError(f"cannot take the address of {type(x)}")
