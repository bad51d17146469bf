"""Host-side mirror of scripts-part2 of the reference (same names, argument meaning, return values and error
behaviour), on top of the C ABI of libb200stencil.so.  Device arrays are torch CUDA float64 tensors laid out exactly
like a Julia CuArray{Float64,2}: shape (nx, ny), column-major (strides (1, nx)) -- use to_device()/to_host().

Mirrored entry points (reference file:line):
  MGOpt, CoarseSolver_t                 scripts-part2/multigrid.jl:10-22
  preallocate_buffers                   scripts-part2/multigrid.jl:25-38      (returns the L1 handle: level table + graph)
  MGsolve_2DPoisson                     scripts-part2/multigrid.jl:41-84
  Vcycle_2DPoisson                      scripts-part2/multigrid.jl:91-170
  iteration_2DPoisson, residual_2DPoisson_wrapper, restrict_wrapper, prolongate_wrapper   :223-258, :344-358, :451-472
  matrix_free_matvec_prod_wrapper, cg   scripts-part2/krylov.jl:37-91
  apply_boundary_conditions[_dirichlet|_neumann], load    scripts-part2/part2_utils.jl:11-39
  SimIn_t, SimOut_t, navier_stokes_2D   scripts-part2/part2.jl:30-55,140-262
"""
import collections
import ctypes as C
import time
import warnings

import numpy as np

from . import _capi as capi

# enums -----------------------------------------------------------------------------------------------------------
jacobi, conjugate_gradient = capi.COARSE_JACOBI, capi.COARSE_CG
serial, parallel, parallel_shmem = capi.POLICY_SERIAL, capi.POLICY_PARALLEL, capi.POLICY_PARALLEL_SHMEM


class MGOpt:
    """multigrid.jl:16-22 (+ the variant-B switches of this library: smoother, restriction)."""

    def __init__(self, coarse_solve_size=5, coarse_solver=jacobi, execution_policy=parallel_shmem,
                 smoother=capi.SMOOTH_JACOBI, restriction=capi.RESTRICT_INJECT, use_graph=True, smem_levels=True,
                 fuse_sweeps=True):
        self.coarse_solve_size = coarse_solve_size
        self.coarse_solver = coarse_solver
        self.execution_policy = execution_policy
        self.smoother = smoother
        self.restriction = restriction
        self.use_graph = use_graph
        self.smem_levels = smem_levels
        self.fuse_sweeps = fuse_sweeps


def _torch():
    import torch
    return torch


def to_device(a, device=0):
    """numpy (nx, ny) -> CUDA tensor of shape (nx, ny) stored column-major like a Julia array."""
    torch = _torch()
    a = np.asarray(a, dtype=np.float64)
    t = torch.from_numpy(np.ascontiguousarray(a.T)).to(f"cuda:{device}")
    return t.T


def zeros(nx, ny, device=0):
    torch = _torch()
    return torch.zeros((ny, nx), dtype=torch.float64, device=f"cuda:{device}").T


def to_host(t):
    return np.asfortranarray(t.detach().cpu().numpy())


def _chk(t):
    torch = _torch()
    assert t.is_cuda and t.dtype == torch.float64 and t.dim() == 2, "need a 2-D CUDA float64 tensor"
    nx, ny = t.shape
    assert t.stride() == (1, nx) or nx == 1 or ny == 1, "array must be column-major (use part2.to_device)"
    return int(nx), int(ny)


def _stream():
    torch = _torch()
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class MGHandle:
    """What preallocate_buffers returns here: the level hierarchy, work arrays and the captured V-cycle graph."""

    def __init__(self, nx, ny, opt=None, device=0):
        opt = opt or MGOpt()
        self._L = capi.lib()
        self.nx, self.ny, self.opt, self.device = nx, ny, opt, device
        cfg = capi.MGConfig(nx, ny, opt.coarse_solve_size, opt.coarse_solver, opt.smoother, opt.restriction, device,
                            int(bool(opt.use_graph)), int(bool(opt.smem_levels)), int(opt.fuse_sweeps))
        self._h = C.c_void_p()
        if opt.execution_policy == serial:
            raise capi.B2SError(capi.ERR_NOT_IMPLEMENTED, "execution policy serial is a CPU-only debug path")
        capi.check(self._L.b2s_mg_create(C.byref(self._h), C.byref(cfg)))

    def close(self):
        if getattr(self, "_h", None):
            self._L.b2s_mg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def solve(self, u, f, h, c, tol, niters, apply_BCs, want_hist=False):
        assert _chk(u) == (self.nx, self.ny) and _chk(f) == (self.nx, self.ny)
        _torch().cuda.current_stream().synchronize()
        r, nc = C.c_double(), C.c_int()
        hist = np.zeros(max(niters, 1)) if want_hist else None
        capi.check(self._L.b2s_mg_solve(self._h, capi.ptr(u), capi.ptr(f), h, c, tol, int(niters), int(bool(apply_BCs)),
                                        C.byref(r), C.byref(nc), capi.ptr(hist) if want_hist else None))
        return (r.value, nc.value, hist[:nc.value]) if want_hist else (r.value, nc.value)

    def vcycle(self, u, rhs, h, c, tol, apply_BCs):
        assert _chk(u) == (self.nx, self.ny) and _chk(rhs) == (self.nx, self.ny)
        _torch().cuda.current_stream().synchronize()
        r = C.c_double()
        capi.check(self._L.b2s_mg_vcycle(self._h, capi.ptr(u), capi.ptr(rhs), h, c, tol, int(bool(apply_BCs)), C.byref(r)))
        return r.value

    def cycles(self, u, f, h, c, tol, ncycles, apply_BCs=False):
        _torch().cuda.current_stream().synchronize()
        r, ms = C.c_double(), C.c_double()
        capi.check(self._L.b2s_mg_cycles(self._h, capi.ptr(u), capi.ptr(f), h, c, tol, int(ncycles), int(bool(apply_BCs)),
                                         C.byref(r), C.byref(ms)))
        return r.value, ms.value

    def pcg(self, u, f, h, c, tol, maxit, tol_mode=capi.PCG_TOL_INITIAL_RESIDUAL):
        """MG-preconditioned CG (extension): returns (r_rms, iterations). tol_mode: relative to the initial residual, or
        MGsolve's criterion r_rms < tol * f_rms (capi.PCG_TOL_RHS)."""
        assert _chk(u) == (self.nx, self.ny) and _chk(f) == (self.nx, self.ny)
        _torch().cuda.current_stream().synchronize()
        r, it = C.c_double(), C.c_int()
        capi.check(self._L.b2s_mg_pcg_solve2(self._h, capi.ptr(u), capi.ptr(f), h, c, tol, int(maxit), int(tol_mode),
                                             C.byref(r), C.byref(it)))
        return r.value, it.value

    def profile_kernels(self, u, f, h, c, reps=20):
        """Average device time (ms) of every kernel of one fused V-cycle in isolation (b2s_mg_profile_kernels):
        [{"kernel": "down"|"up"|"tail", "level": l, "grid": [nx, ny], "ms": t}, ...]."""
        _torch().cuda.current_stream().synchronize()
        nl = C.c_int()
        md, mu = (C.c_double * 24)(), (C.c_double * 24)()
        mt = C.c_double()
        lx, ly = (C.c_int * 24)(), (C.c_int * 24)()
        capi.check(self._L.b2s_mg_profile_kernels(self._h, capi.ptr(u), capi.ptr(f), h, c, int(reps), C.byref(nl), md, mu,
                                                  C.byref(mt), lx, ly))
        out = []
        for l in range(nl.value):
            out.append({"kernel": "down", "level": l, "grid": [lx[l], ly[l]], "ms": md[l]})
            out.append({"kernel": "up", "level": l, "grid": [lx[l], ly[l]], "ms": mu[l]})
        out.append({"kernel": "tail", "level": nl.value, "grid": None, "ms": mt.value})
        return out

    def last_coarse_sweeps(self):
        n = C.c_int()
        capi.check(self._L.b2s_mg_last_coarse_sweeps(self._h, C.byref(n)))
        return n.value

    def stats(self):
        n, ms = C.c_longlong(), C.c_double()
        capi.check(self._L.b2s_mg_stats(self._h, C.byref(n), C.byref(ms)))
        return n.value, ms.value


def preallocate_buffers(nx, ny, opt=None, device=0):
    return MGHandle(nx, ny, opt, device)


def MGsolve_2DPoisson(u, f, h, c, tol, niters, apply_BCs, *, opt=None, verbose=False, prealloc_dict=None,
                      return_cycles=False):
    """r_rms = MGsolve_2DPoisson!(u, f, h, c, tol, niters, apply_BCs; opt, verbose, prealloc_dict)   multigrid.jl:41."""
    nx, ny = _chk(u)
    opt = opt or (prealloc_dict.opt if prealloc_dict is not None else MGOpt())
    own = prealloc_dict is None
    hd = prealloc_dict if not own else MGHandle(nx, ny, opt, u.device.index or 0)  # asserts :45-46 -> B2SError
    try:
        r, nc, hist = hd.solve(u, f, h, c, tol, niters, apply_BCs, want_hist=True)
        if verbose:
            for i, v in enumerate(hist):
                print(f"Vcycle iter {i + 1}: r_rms / f_rms = {v}")
        if nc == niters and len(hist) and not (hist[-1] < tol):  # :78-80 -- a warning, not an error
            warnings.warn(f"MGsolve_2DPoisson did not converge to {tol} within {niters} V-cycles")
        return (r, nc) if return_cycles else r
    finally:
        if own:
            hd.close()


def Vcycle_2DPoisson(u_f, rhs, h, c, tol, coarse_solve_size, coarse_solver, execution_policy, apply_BCs, *,
                     prealloc_dict=None):
    """res_rms = Vcycle_2DPoisson!(...)   multigrid.jl:91-170."""
    nx, ny = _chk(u_f)
    if ((nx - 1) % 2) or ((ny - 1) % 2):
        raise capi.B2SError(capi.ERR_BAD_SIZE, "ERROR:not a power of 2")  # :95-97
    own = prealloc_dict is None
    hd = prealloc_dict if not own else MGHandle(nx, ny, MGOpt(coarse_solve_size, coarse_solver, execution_policy),
                                                u_f.device.index or 0)
    try:
        return hd.vcycle(u_f, rhs, h, c, tol, apply_BCs)
    finally:
        if own:
            hd.close()


def residual_2DPoisson_wrapper(u, f, h, c, res, execution_policy=parallel_shmem):
    nx, ny = _chk(u)
    capi.check(capi.lib().b2s_residual2d(capi.ptr(u), capi.ptr(f), h, c, capi.ptr(res), nx, ny, execution_policy, _stream()))


def iteration_2DPoisson(u, f, h, c, res, execution_policy=parallel_shmem, *, alpha=4.0 / 5.0):
    nx, ny = _chk(u)
    r = C.c_double()
    capi.check(capi.lib().b2s_iteration2d(capi.ptr(u), capi.ptr(f), h, c, capi.ptr(res), nx, ny, alpha, execution_policy,
                                          C.byref(r), _stream()))
    return r.value


def rbgs_2DPoisson(u, f, h, c):
    nx, ny = _chk(u)
    r = C.c_double()
    capi.check(capi.lib().b2s_rbgs2d(capi.ptr(u), capi.ptr(f), h, c, nx, ny, C.byref(r), _stream()))
    return r.value


def restrict_wrapper(fine, coarse, apply_BCs, execution_policy=parallel_shmem, full_weighting=False):
    nx, ny = _chk(fine)
    assert _chk(coarse) == (1 + (nx - 1) // 2, 1 + (ny - 1) // 2)
    fn = capi.lib().b2s_restrict_fw2d if full_weighting else capi.lib().b2s_restrict_inject2d
    capi.check(fn(capi.ptr(fine), capi.ptr(coarse), nx, ny, int(bool(apply_BCs)), _stream()))


def prolongate_wrapper(coarse, fine, apply_BCs, execution_policy=parallel_shmem):
    nx, ny = _chk(fine)
    assert _chk(coarse) == (1 + (nx - 1) // 2, 1 + (ny - 1) // 2)
    capi.check(capi.lib().b2s_prolongate2d(capi.ptr(coarse), capi.ptr(fine), nx, ny, int(bool(apply_BCs)), _stream()))


def matrix_free_matvec_prod_wrapper(p, hx, hy, c, p_hat, execution_policy=parallel_shmem):
    nx, ny = _chk(p)
    capi.check(capi.lib().b2s_matvec2d(capi.ptr(p), hx, hy, c, capi.ptr(p_hat), nx, ny, execution_policy, _stream()))
    _torch().cuda.current_stream().synchronize()  # @synchronize() krylov.jl:51


def cg(x_in, b, hx, hy, c, tol, Nmax, *, execution_policy=parallel_shmem, verbose=False, return_iters=False):
    """res_rms = cg!(x_in, b, hx, hy, c, tol, Nmax; execution_policy, verbose)   krylov.jl:55-91."""
    nx, ny = _chk(x_in)
    r, it = C.c_double(), C.c_int()
    capi.check(capi.lib().b2s_cg_solve(capi.ptr(x_in), capi.ptr(b), hx, hy, c, tol, int(Nmax), nx, ny, execution_policy,
                                       C.byref(r), C.byref(it), _stream()))
    return (r.value, it.value) if return_iters else r.value


def apply_boundary_conditions(T):
    nx, ny = _chk(T)
    capi.check(capi.lib().b2s_apply_bc2d(capi.ptr(T), nx, ny, 0, _stream()))


def apply_boundary_conditions_dirichlet(T):
    nx, ny = _chk(T)
    capi.check(capi.lib().b2s_apply_bc2d(capi.ptr(T), nx, ny, 1, _stream()))


def apply_boundary_conditions_neumann(T):
    nx, ny = _chk(T)
    capi.check(capi.lib().b2s_apply_bc2d(capi.ptr(T), nx, ny, 2, _stream()))


def load(path):
    """part2_utils.jl:11-19: Int32 nx, Int32 ny, nx*ny Float64 column-major."""
    with open(path, "rb") as f:
        nx, ny = np.fromfile(f, dtype=np.int32, count=2)
        a = np.fromfile(f, dtype=np.float64, count=int(nx) * int(ny))
    return np.asfortranarray(a.reshape((int(nx), int(ny)), order="F"))


def dot(x, y):
    r = C.c_double()
    capi.check(capi.lib().b2s_dot(capi.ptr(x), capi.ptr(y), x.numel(), C.byref(r), _stream()))
    return r.value


def sumsq(x):
    r = C.c_double()
    capi.check(capi.lib().b2s_sumsq(capi.ptr(x), x.numel(), C.byref(r), _stream()))
    return r.value


# ---- Navier-Stokes driver ---------------------------------------------------------------------------------------
class SimIn_t:
    """part2.jl:30-46 (defaults of the inner constructor :45)."""

    def __init__(self, **kw):
        self.k, self.Ra, self.Pr = 1.0, 1.0e6, 1.0e-3
        self.nx, self.ny = 257, 65
        self.ttot, self.beta, self.niters, self.tol = 0.1, 0.0, 50, 1.0e-3
        self.a_dif, self.a_adv = 0.15, 0.4
        self.T_init_strategy, self.W_init_strategy = "cosine", "random"
        self.W_init = None  # array for W_init_strategy == "W_from_file" / explicit data
        self.seed = 0
        for k, v in kw.items():
            setattr(self, k, v)


SimOut_t = collections.namedtuple("SimOut_t", "T W S t_elapsed timed_iters")


class NavierStokes2D:
    def __init__(self, opt, mgopt=None, device=0, solver=capi.NS_SOLVER_VCYCLE):
        """solver: NS_SOLVER_VCYCLE (the reference: plain V-cycle iteration) or NS_SOLVER_MG_PCG (MG-preconditioned CG for
        the S and W solves; needs mgopt.restriction = full weighting)."""
        self._L = capi.lib()
        mgopt = mgopt or MGOpt()
        p = capi.NS2DParams(opt.k, opt.Ra, opt.Pr, opt.nx, opt.ny, opt.ttot, opt.beta, opt.niters, opt.tol, opt.a_dif,
                            opt.a_adv)
        cfg = capi.MGConfig(opt.nx, opt.ny, mgopt.coarse_solve_size, mgopt.coarse_solver, mgopt.smoother,
                            mgopt.restriction, device, int(bool(mgopt.use_graph)), int(bool(mgopt.smem_levels)),
                            int(mgopt.fuse_sweeps))
        self.shape = (opt.nx, opt.ny)
        self._h = C.c_void_p()
        capi.check(self._L.b2s_ns2d_create(C.byref(self._h), C.byref(p), C.byref(cfg)))
        if solver != capi.NS_SOLVER_VCYCLE:
            capi.check(self._L.b2s_ns2d_set_solver(self._h, int(solver)))

    def close(self):
        if getattr(self, "_h", None):
            self._L.b2s_ns2d_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_field(self, which, a):
        a = np.asfortranarray(a, dtype=np.float64)
        assert a.shape == self.shape
        capi.check(self._L.b2s_ns2d_set_field(self._h, "TWS".index(which), capi.ptr(a)))

    def get_field(self, which):
        a = np.zeros(self.shape, order="F")
        capi.check(self._L.b2s_ns2d_get_field(self._h, "TWS".index(which), capi.ptr(a)))
        return a

    def get_aux(self, name):
        a = np.zeros(self.shape, order="F")
        capi.check(self._L.b2s_ns2d_get_aux(self._h, ["vx", "vy", "Ra_dTdx", "dT2", "dW2"].index(name), capi.ptr(a)))
        return a

    def init_cosine(self, which):
        capi.check(self._L.b2s_ns2d_init_cosine(self._h, "TWS".index(which)))

    def step(self):
        info = capi.NS2DStepInfo()
        capi.check(self._L.b2s_ns2d_step(self._h, C.byref(info)))
        return info


def navier_stokes_2D(*, opt=None, verbose=True, do_vis=False, testmode=False, mgopt=None, device=0, return_infos=False,
                     solver=capi.NS_SOLVER_VCYCLE):
    """Drop-in for part2.jl:140-262. Returns SimOut_t(T, W, S, t_elapsed, timed_iters) with host arrays.
    `solver` / `mgopt` select MG-preconditioned CG for the S and W solves (extension; default: the reference's cycling)."""
    opt = opt or SimIn_t()
    sim = NavierStokes2D(opt, mgopt, device, solver)
    try:
        if opt.T_init_strategy == "cosine":
            sim.init_cosine("T")
        else:
            sim.set_field("T", np.random.default_rng(opt.seed).random((opt.nx, opt.ny)))
        if opt.W_init is not None:
            sim.set_field("W", opt.W_init)
        elif opt.W_init_strategy == "cosine":
            sim.init_cosine("W")
        else:
            sim.set_field("W", np.random.default_rng(opt.seed + 1).random((opt.nx, opt.ny)))
        tic, sim_time, step, infos = 0.0, 0.0, 0, []
        while sim_time < opt.ttot:
            if step == 3:
                tic = time.time()
            info = sim.step()
            infos.append((info.dt, info.cycles_S, info.cycles_T, info.cycles_W, info.r_S, info.r_T, info.r_W))
            sim_time += info.dt
            step += 1
            if (step - 1) % 20 == 0 and verbose:
                print(f"time, step: {sim_time} {step}")
            if testmode:
                break
        out = SimOut_t(sim.get_field("T"), sim.get_field("W"), sim.get_field("S"), time.time() - tic, step - 3)
        if verbose:
            print(f"time, step: {sim_time} {step}")
        return (out, infos) if return_infos else out
    finally:
        sim.close()


# ---- benchmark / smoke helpers ------------------------------------------------------------------------------------
def mg_algorithmic_bytes(nx, ny, coarse_solve_size=5):
    """132 B x sum over non-coarsest levels of N_l (4 sweeps x 3 arrays + residual 2 + coarse rhs 1/4 +
    prolong/correct 2 1/4 = 16.5 doubles per point)."""
    tot, lx, ly = 0, nx, ny
    while min(lx, ly) > coarse_solve_size:
        tot += lx * ly
        lx, ly = 1 + (lx - 1) // 2, 1 + (ly - 1) // 2
    return 132.0 * tot


def mg_fused_min_bytes(nx, ny, coarse_solve_size=5):
    """Compulsory traffic of the fused formulation (2 kernels per level): down reads u, f and writes u_s, rc/4, ec/4
    (3.5 doubles per point), up reads u_s, f, ec/4 and writes u (3.25) = 54 B per point of every non-coarsest level."""
    return mg_algorithmic_bytes(nx, ny, coarse_solve_size) / 132.0 * 54.0


def bench_vcycle(device=0, hbm_peak_gbs=6527.8, sizes=(1025, 2049, 4097), ncycles=50, seed=1, opt=None, e2e=True):
    """Config #2 (multigrid_bench.jl shape): x = 0, b ~ U[0,1) on all entries, c = 0, tol 1e-6; DoF/s per V-cycle with
    fields resident on the device (CUDA events inside the library), plus the whole-solve time and cycle count."""
    torch = _torch()
    out = {"unit": "DoF/s per V-cycle", "sizes": {}}
    for n in sizes:
        h = 1.0 / (n - 1)
        b = to_device(np.random.default_rng(seed).random((n, n)), device)
        hd = MGHandle(n, n, opt if opt is not None else MGOpt(), device)
        x = zeros(n, n, device)
        hd.cycles(x, b, h, 0.0, 1e-6, 3)  # warm-up (graph instantiation)
        hd.cycles(x, b, h, 0.0, 1e-6, 300)  # ... and let the clocks settle (the cycle is latency-bound: ~30 ms of work)
        times = []
        for _ in range(5):  # median of 5 timings of `ncycles` V-cycles each (CUDA events inside the library)
            x.zero_()
            torch.cuda.synchronize()
            l0, _ = hd.stats()
            _, ms_i = hd.cycles(x, b, h, 0.0, 1e-6, ncycles)
            l1, _ = hd.stats()
            times.append(ms_i)
        ms = sorted(times)[len(times) // 2]
        solve_s = float("inf")
        for _ in range(3):  # whole MGsolve (x = 0 -> 1e-6), wall clock, best of 3 (the first one also sizes the batches)
            x.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r, nc = hd.solve(x, b, h, 0.0, 1e-6, 100, False)
            torch.cuda.synchronize()
            solve_s = min(solve_s, time.perf_counter() - t0)
        # end to end through the public entry point with HOST arrays: upload of the right-hand side from pinned memory,
        # MGsolve_2DPoisson (x = 0 initial guess created on the device like the reference's CUDA.zeros), download of x
        b_host = torch.from_numpy(np.ascontiguousarray(np.random.default_rng(seed).random((n, n)).T)).pin_memory()
        x_host = torch.empty((n, n), dtype=torch.float64).pin_memory()
        e2e_s = float("inf")
        for _ in range(3 if e2e else 0):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            bd = b_host.to(f"cuda:{device}", non_blocking=True).T
            xd = zeros(n, n, device)
            hd.solve(xd, bd, h, 0.0, 1e-6, 100, False)
            x_host.copy_(xd.T, non_blocking=True)
            torch.cuda.synchronize()
            e2e_s = min(e2e_s, time.perf_counter() - t0)
        per = ms / ncycles * 1e-3
        ab = mg_algorithmic_bytes(n, n)
        # Per kernel, in isolation (CUDA events, 20 back-to-back launches): the fused formulation's compulsory traffic is
        # 28 B per point of the level for the downward kernel (read u, f; write u_s, rc/4, ec/4) and 26 B for the upward
        # one (read u_s, f, ec/4; write u) -- THE byte model of every fraction below; levels whose arrays fit the 126 MB
        # L2 (<= 2049^2) are served from L2, so their fraction of the HBM peak is a lower bound on what limits them.
        kernels, dominant = [], None
        try:
            x.zero_()
            for k in hd.profile_kernels(x, b, h, 0.0):
                if k["kernel"] == "tail":
                    k.update(name="mg_mid_cluster_kernel / mg_coarse_kernel (all levels below, resident in shared memory)",
                             algorithmic_bytes=0.0, achieved_gbs=None, frac_of_hbm_peak=None, bound="latency")
                else:
                    pts = float(k["grid"][0]) * k["grid"][1]
                    by = (28.0 if k["kernel"] == "down" else 26.0) * pts
                    streaming = pts > 1.5e6 and (opt is None or opt.smoother == 0)
                    k.update(name=("mg_%s_stream2_kernel" if streaming else "mg_%s_kernel<tile>") % k["kernel"]
                             if (opt is None or opt.smoother == 0) else "mg_%s_rb_kernel<tile>" % k["kernel"],
                             algorithmic_bytes=by, achieved_gbs=by / (k["ms"] * 1e-3) / 1e9,
                             frac_of_hbm_peak=by / (k["ms"] * 1e-3) / 1e9 / hbm_peak_gbs,
                             bound="hbm" if pts * 8 * 3 > 126e6 else "l2-resident / latency")
                kernels.append(k)
            dominant = max(kernels, key=lambda k: k["ms"])
        except Exception as e:  # pragma: no cover
            kernels = [{"unavailable": f"{type(e).__name__}: {e}"}]
        out["sizes"][str(n)] = {"dof_per_s": n * n / per, "ms_per_vcycle": per * 1e3,
                                "ms_per_vcycle_min_max_of_5": [min(times) / ncycles, max(times) / ncycles],
                                "vcycles_to_1e-6": nc,
                                "solve_ms": solve_s * 1e3,
                                "e2e": None if not e2e else {
                                    "solve_ms": e2e_s * 1e3, "dof_per_s_per_vcycle": n * n * nc / e2e_s,
                                    "h2d_bytes": n * n * 8, "d2h_bytes": n * n * 8,
                                    "what": "pinned host rhs -> device, MGsolve to 1e-6, solution -> pinned host (best of 3)"},
                                # one byte model for every fraction: the fused formulation's compulsory 54 B per point of
                                # every non-coarsest level (28 down + 26 up)
                                "roofline": {"bound": "hbm" if n * n * 8 * 3 > 126e6 else "latency (arrays are L2-resident)",
                                             "model": "fused: 54 B per point of every non-coarsest level per V-cycle",
                                             "algorithmic_bytes_per_vcycle": mg_fused_min_bytes(n, n),
                                             "achieved": mg_fused_min_bytes(n, n) / per / 1e9, "peak": hbm_peak_gbs,
                                             "unit": "GB/s", "frac": mg_fused_min_bytes(n, n) / per / 1e9 / hbm_peak_gbs,
                                             "dominant_kernel": dominant, "kernels": kernels,
                                             "sum_of_isolated_kernels_ms": sum(k.get("ms", 0.0) for k in kernels)},
                                # SURVEY 8d's un-fused accounting (132 B per point: one pass per sweep). The fused kernels do
                                # not make those passes, so this is an EFFECTIVE rate for comparison with un-fused
                                # implementations, not a fraction of any hardware peak.
                                "survey_model_bytes_per_vcycle": ab, "survey_model_effective_gbs": ab / per / 1e9,
                                "kernel_launches_per_vcycle": (l1 - l0) / ncycles}
        hd.close()
    return out


def bench_config2_matrix(device=0, n=1025, ncycles=30):
    """Config #2 as multigrid_bench.jl sweeps it: coarse_solve_size in {5, 9} x coarse solver {Jacobi, CG} for variant A
    (damped Jacobi + injection) and variant B (red-black Gauss-Seidel + full weighting): ms per V-cycle, V-cycles to 1e-6."""
    rows = []
    for variant in ("A", "B"):
        for cs in (5, 9):
            for solver in (jacobi, conjugate_gradient):
                opt = MGOpt(coarse_solve_size=cs, coarse_solver=solver, smoother=1 if variant == "B" else 0,
                            restriction=1 if variant == "B" else 0)
                d = bench_vcycle(device=device, sizes=(n,), ncycles=ncycles, opt=opt, e2e=False)["sizes"][str(n)]
                rows.append({"variant": variant, "coarse_solve_size": cs, "coarse_solver": "jacobi" if solver == jacobi else "cg",
                             "ms_per_vcycle": d["ms_per_vcycle"], "dof_per_s": d["dof_per_s"], "vcycles_to_1e-6": d["vcycles_to_1e-6"],
                             "solve_ms": d["solve_ms"]})
    return rows


def bench_navier_stokes(device=0, n=2049, steps=8, seed=1, solver=capi.NS_SOLVER_VCYCLE, mgopt=None):
    """Config #4: 2-D streamfunction-vorticity Navier-Stokes, semi-implicit (beta = 0.5, Pr = 0.1, tol 1e-7: the
    published experiment's settings, part2_semi_implicit_vs_explicit_experiments.jl:38-44) on an n x n grid, W ~ U[0,1).
    Times the steps from the 4th on (the reference starts its timer at step 3, part2.jl:182-184). `solver`: plain V-cycle
    iteration like the reference, or MG-preconditioned CG for the S and W solves (cycle-equivalents = CG iterations)."""
    torch = _torch()
    opt = SimIn_t(nx=n, ny=n, beta=0.5, Pr=0.1, tol=1.0e-7, niters=50)
    sim = NavierStokes2D(opt, mgopt or MGOpt(), device, solver)
    sim.init_cosine("T")
    sim.set_field("W", np.random.default_rng(seed).random((n, n)))
    infos, t0 = [], None
    for step in range(steps):
        if step == 3:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
        i = sim.step()
        infos.append((i.dt, i.cycles_S, i.cycles_T, i.cycles_W))
    torch.cuda.synchronize()
    dt_wall = time.perf_counter() - t0
    timed = infos[3:]
    cyc = sum(a + b + c for _, a, b, c in timed)
    sim.close()
    return {"grid": [n, n], "beta": 0.5, "Pr": 0.1, "tol": 1e-7, "timed_steps": len(timed),
            "solver": "mg_pcg (S, W) + v-cycles (T)" if solver == capi.NS_SOLVER_MG_PCG else "v-cycles (reference)",
            "ms_per_step": dt_wall / len(timed) * 1e3, "vcycles_per_step": cyc / len(timed),
            "cycles_S_T_W": [list(x[1:]) for x in infos], "dof_per_s_per_vcycle": n * n * cyc / dt_wall,
            "note": "whole time step incl. 3 solves, velocity/dt reduction, fused stencil + rhs kernel; wall clock"}


def smoke_check(O):
    """One small multigrid solve on cuda:0 against the oracle (used by __graft_entry__.smoke())."""
    n = 129
    h = 1.0 / (n - 1)
    b = np.asfortranarray(np.random.default_rng(1).random((n, n)))
    xo = O.farray((n, n))
    ro, nco, _ = O.mgsolve2d(xo, b, h, 0.0, 1e-6, 100)
    x = zeros(n, n)
    r, nc = MGsolve_2DPoisson(x, to_device(b), h, 0.0, 1e-6, 100, False, return_cycles=True)
    assert nc == nco == 7, (nc, nco)
    err = np.max(np.abs(to_host(x) - xo)) / np.max(np.abs(xo))
    assert err < 1e-12, err
    assert abs(r - ro) <= 1e-9 * ro
