"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

NumPy arrays are Fortran-ordered (column-major, x fastest) exactly like the reference's Julia arrays.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

HALO_REFERENCE_LAG2, HALO_CONSISTENT = 0, 1
BC_LITERAL, BC_PROPER = 0, 1
COARSE_JACOBI, COARSE_CG = 0, 1
SMOOTH_JACOBI, SMOOTH_RBGS = 0, 1
RESTRICT_INJECT, RESTRICT_FW = 0, 1

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class MGOpt(C.Structure):
    _fields_ = [("coarse_solve_size", C.c_int), ("coarse_solver", C.c_int), ("smoother", C.c_int),
                ("restriction", C.c_int), ("unfused", C.c_int)]

    def __init__(self, coarse_solve_size=5, coarse_solver=COARSE_JACOBI, smoother=SMOOTH_JACOBI,
                 restriction=RESTRICT_INJECT, unfused=0):
        super().__init__(coarse_solve_size, coarse_solver, smoother, restriction, unfused)


class NSParams(C.Structure):
    _fields_ = [("k", C.c_double), ("Ra", C.c_double), ("Pr", C.c_double), ("nx", C.c_int), ("ny", C.c_int),
                ("ttot", C.c_double), ("beta", C.c_double), ("niters", C.c_int), ("tol", C.c_double),
                ("a_dif", C.c_double), ("a_adv", C.c_double)]

    def __init__(self, **kw):
        # SimIn_t defaults, scripts-part2/part2.jl:45
        d = dict(k=1.0, Ra=1.0e6, Pr=1.0e-3, nx=257, ny=65, ttot=0.1, beta=0.0, niters=50, tol=1.0e-3,
                 a_dif=0.15, a_adv=0.4)
        d.update(kw)
        super().__init__(**d)


class NSStepInfo(C.Structure):
    _fields_ = [("dt", C.c_double), ("cycles_S", C.c_int), ("cycles_T", C.c_int), ("cycles_W", C.c_int),
                ("r_S", C.c_double), ("r_T", C.c_double), ("r_W", C.c_double)]


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("diffusion3d_oracle.c", "multigrid2d_oracle.c")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "PORTABLE=1", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.orc_diff3d_create.restype = C.c_void_p
        L.orc_diff3d_create.argtypes = [C.c_int] * 10
        L.orc_diff3d_destroy.argtypes = [C.c_void_p]
        L.orc_diff3d_iterate_once.restype = C.c_double
        L.orc_diff3d_iterate_once.argtypes = [C.c_void_p]
        L.orc_diff3d_iterate.argtypes = [C.c_void_p, C.c_int, _dp]
        L.orc_diff3d_solve_timestep.restype = C.c_int
        L.orc_diff3d_solve_timestep.argtypes = [C.c_void_p, C.c_double, C.c_int, _dp]
        L.orc_diff3d_advance_time.argtypes = [C.c_void_p]
        L.orc_diff3d_get.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp]
        L.orc_diff3d_gather.argtypes = [C.c_void_p, _dp]
        L.orc_diff3d_params.argtypes = [C.c_void_p, _dp]
        L.orc_diff3d_set_array.argtypes = [C.c_void_p]
        L.orc_diff3d_set_array.restype = None
        L.orc_diff3d_num_timesteps.restype = C.c_int
        L.orc_diff3d_num_timesteps.argtypes = [C.c_double, C.c_double]
        L.orc_diff3d_run.restype = C.c_long
        L.orc_diff3d_run.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int, _ip]
        L.orc_num_threads.restype = C.c_int
        L.orc_mg_last_coarse_sweeps.restype = C.c_long
        for name in ("orc_bc_dirichlet", "orc_bc_neumann", "orc_bc_apply"):
            getattr(L, name).argtypes = [_dp, C.c_int, C.c_int]
        L.orc_residual2d.argtypes = [_dp, _dp, C.c_double, C.c_double, _dp, C.c_int, C.c_int]
        L.orc_sumsq.restype = C.c_double
        L.orc_sumsq.argtypes = [_dp, C.c_int, C.c_int]
        L.orc_jacobi2d.restype = C.c_double
        L.orc_jacobi2d.argtypes = [_dp, _dp, C.c_double, C.c_double, _dp, C.c_int, C.c_int, C.c_double, C.c_int]
        L.orc_gs2d_lex.restype = C.c_double
        L.orc_gs2d_lex.argtypes = [_dp, _dp, C.c_double, C.c_double, C.c_int, C.c_int, C.c_double]
        L.orc_rbgs2d.restype = C.c_double
        L.orc_rbgs2d.argtypes = [_dp, _dp, C.c_double, C.c_double, C.c_int, C.c_int]
        for name in ("orc_restrict_inject", "orc_restrict_fw", "orc_prolongate"):
            getattr(L, name).argtypes = [_dp, _dp, C.c_int, C.c_int, C.c_int]
        L.orc_matvec2d.argtypes = [_dp, C.c_double, C.c_double, C.c_double, _dp, C.c_int, C.c_int]
        L.orc_cg2d.restype = C.c_double
        L.orc_cg2d.argtypes = [_dp, _dp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int,
                               _ip]
        L.orc_vcycle2d.restype = C.c_double
        L.orc_vcycle2d.argtypes = [_dp, _dp, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int,
                                   C.POINTER(MGOpt)]
        L.orc_mgsolve2d.restype = C.c_double
        L.orc_mgsolve2d.argtypes = [_dp, _dp, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.POINTER(MGOpt), _ip, _dp]
        L.orc_mg_pcg2d.restype = C.c_double
        L.orc_mg_pcg2d.argtypes = [_dp, _dp, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.POINTER(MGOpt), _ip]
        L.orc_ns_set_solver.argtypes = [C.c_int]
        L.orc_ns_set_solver.restype = None
        L.orc_ns_init_cosine.argtypes = [_dp, C.c_int, C.c_int]
        L.orc_ns_step.argtypes = [C.POINTER(NSParams), C.POINTER(MGOpt), _dp, _dp, _dp, C.POINTER(NSStepInfo), _dp]
        _LIB = L
    return _LIB


def _p(a):
    assert a.dtype == np.float64 and a.flags.f_contiguous, "oracle arrays are float64, column-major"
    return a.ctypes.data_as(_dp)


def farray(shape):
    return np.zeros(shape, dtype=np.float64, order="F")


def load_bin(path):
    """scripts-part2/part2_utils.jl:11-19: Int32 nx, Int32 ny, nx*ny Float64 column-major."""
    with open(path, "rb") as f:
        nx, ny = np.fromfile(f, dtype=np.int32, count=2)
        a = np.fromfile(f, dtype=np.float64, count=int(nx) * int(ny))
    return np.asfortranarray(a.reshape((int(nx), int(ny)), order="F"))


class Diffusion3D:
    """Emulated-rank oracle of diffusion_3D_kernel_programming (scripts-part1/part1_kernel_programming.jl:99)."""

    def __init__(self, nx, ny, nz, dims=(1, 1, 1), halo_mode=HALO_REFERENCE_LAG2, bc_mode=BC_LITERAL,
                 scale_physical_size=False, unfused_norm=False, array=False):
        """array=True: diffusion_3D_array_programming (scripts-part1/part1_array_programming.jl:20) instead."""
        self.L = lib()
        self.n = (nx, ny, nz)
        self.dims = tuple(dims)
        self.h = self.L.orc_diff3d_create(nx, ny, nz, dims[0], dims[1], dims[2], halo_mode, bc_mode,
                                          int(scale_physical_size), int(unfused_norm))
        if array:
            self.L.orc_diff3d_set_array(self.h)
        p = np.zeros(8)
        self.L.orc_diff3d_params(self.h, p.ctypes.data_as(_dp))
        self.dx, self.dy, self.dz, self.dt, self.dtau, self.lx, self.ly, self.lz = p

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_diff3d_destroy(self.h)
            self.h = None

    def iterate(self, n):
        e = np.zeros(n)
        self.L.orc_diff3d_iterate(self.h, n, e.ctypes.data_as(_dp))
        return e

    def solve_timestep(self, tol, iter_max=100000):
        err = C.c_double()
        it = self.L.orc_diff3d_solve_timestep(self.h, tol, iter_max, C.byref(err))
        return it, err.value

    def advance_time(self):
        self.L.orc_diff3d_advance_time(self.h)

    def get(self, which, rank=0):
        a = farray(self.n)
        self.L.orc_diff3d_get(self.h, rank, {"Ht": 0, "Htau": 1, "Htau2": 2, "residual": 3}[which], _p(a))
        return a

    def gather(self):
        a = farray(tuple(n * d for n, d in zip(self.n, self.dims)))
        self.L.orc_diff3d_gather(self.h, _p(a))
        return a

    def run(self, ttot=1.0, tol=1e-8, iter_max=100000):
        nt = self.L.orc_diff3d_num_timesteps(ttot, self.dt)
        its = (C.c_int * nt)()
        self.L.orc_diff3d_run(self.h, ttot, tol, iter_max, its)
        return list(its)


def residual2d(u, f, h, c):
    res = farray(u.shape)
    lib().orc_residual2d(_p(u), _p(f), h, c, _p(res), *u.shape)
    return res


def jacobi2d(u, f, h, c, res=None, alpha=0.8, unfused=False):
    res = farray(u.shape) if res is None else res
    return lib().orc_jacobi2d(_p(u), _p(f), h, c, _p(res), u.shape[0], u.shape[1], alpha, int(unfused))


def gs2d_lex(u, f, h, c, alpha=1.0):
    return lib().orc_gs2d_lex(_p(u), _p(f), h, c, u.shape[0], u.shape[1], alpha)


def rbgs2d(u, f, h, c):
    return lib().orc_rbgs2d(_p(u), _p(f), h, c, u.shape[0], u.shape[1])


def _coarse_shape(shape):
    return (1 + (shape[0] - 1) // 2, 1 + (shape[1] - 1) // 2)


def restrict_inject(fine, apply_BCs=False):
    coarse = farray(_coarse_shape(fine.shape))
    lib().orc_restrict_inject(_p(fine), _p(coarse), fine.shape[0], fine.shape[1], int(apply_BCs))
    return coarse


def restrict_fw(fine, apply_BCs=False):
    coarse = farray(_coarse_shape(fine.shape))
    lib().orc_restrict_fw(_p(fine), _p(coarse), fine.shape[0], fine.shape[1], int(apply_BCs))
    return coarse


def prolongate(coarse, fine_shape, apply_BCs=False):
    fine = farray(fine_shape)
    lib().orc_prolongate(_p(coarse), _p(fine), fine_shape[0], fine_shape[1], int(apply_BCs))
    return fine


def matvec2d(T, hx, hy, c, out=None):
    out = farray(T.shape) if out is None else out
    lib().orc_matvec2d(_p(T), hx, hy, c, _p(out), *T.shape)
    return out


def cg2d(x, b, hx, hy, c, tol, Nmax):
    it = C.c_int()
    r = lib().orc_cg2d(_p(x), _p(b), hx, hy, c, tol, Nmax, b.shape[0], b.shape[1], C.byref(it))
    return r, it.value


def vcycle2d(u, rhs, h, c, tol, apply_BCs=False, opt=None):
    opt = opt or MGOpt()
    return lib().orc_vcycle2d(_p(u), _p(rhs), h, c, tol, u.shape[0], u.shape[1], int(apply_BCs), C.byref(opt))


def mgsolve2d(u, f, h, c, tol, niters, apply_BCs=False, opt=None):
    """Returns (r_rms, ncycles, rel_hist)."""
    opt = opt or MGOpt()
    nc = C.c_int()
    hist = np.zeros(niters)
    r = lib().orc_mgsolve2d(_p(u), _p(f), h, c, tol, niters, int(apply_BCs), u.shape[0], u.shape[1], C.byref(opt),
                            C.byref(nc), hist.ctypes.data_as(_dp))
    return r, nc.value, hist[:nc.value]


def mg_pcg2d(u, f, h, c, tol, maxit, opt=None):
    """Returns (r_rms, iterations)."""
    opt = opt or MGOpt()
    it = C.c_int()
    r = lib().orc_mg_pcg2d(_p(u), _p(f), h, c, tol, maxit, u.shape[0], u.shape[1], C.byref(opt), C.byref(it))
    return r, it.value


def ns_step(params, S, T, W, opt=None, want_aux=False, solver=0):
    """solver: 0 plain V-cycle iteration (the reference), 1 MG-preconditioned CG for the S and W solves (extension)."""
    opt = opt or MGOpt()
    info = NSStepInfo()
    lib().orc_ns_set_solver(int(solver))
    aux = np.zeros(7 * S.size) if want_aux else None
    lib().orc_ns_step(C.byref(params), C.byref(opt), _p(S), _p(T), _p(W), C.byref(info),
                      aux.ctypes.data_as(_dp) if want_aux else None)
    if want_aux:
        names = ["vx", "vy", "v", "Ra_dTdx", "dT2", "dW2"]
        aux = {n: np.asfortranarray(aux[i * S.size:(i + 1) * S.size].reshape(S.shape, order="F"))
               for i, n in enumerate(names)}
    return info, aux


def ns_init_cosine(nx, ny):
    M = farray((nx, ny))
    lib().orc_ns_init_cosine(_p(M), nx, ny)
    return M


def num_threads():
    return lib().orc_num_threads()
