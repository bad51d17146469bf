"""Host-side mirror of scripts-part1 of the reference (same names, argument meaning and return values), on top of
the C ABI of libb200stencil.so.  The reference's host language is Julia, which is absent from this image; the Julia
module (julia/B200Stencil.jl) binds the same symbols with ccall and has the same structure as this file.

Mirrored entry points (reference file:line):
  diffusion_3D_kernel_programming   scripts-part1/part1_kernel_programming.jl:99-228
  diffusion_3D_array_programming    scripts-part1/part1_array_programming.jl:20-92  (B2S_ARITH_ARRAY: its own arithmetic,
                                    in-place update of Htau, consistent halos)
  BenchResults                      scripts-part1/part1_kernel_programming.jl:22-29
  main (CLI)                        scripts-part1/part1.jl:25-60
"""
import collections
import ctypes as C
import time

import numpy as np

from . import _capi as capi

BenchResults = collections.namedtuple("BenchResults", "dt Work Performance Memory Intensity Throughput")


class Diffusion3D:
    """L1 solver handle (b2s_diff3d_*): z-slab stack of local grids, one slab per reference MPI rank.

    In-process use: devices = one CUDA ordinal per slab (may repeat). One process per GPU (torch.distributed):
    slab_begin = rank, slab_count = 1, then connect() with the all-gathered IPC blobs.
    """

    def __init__(self, nx, ny, nz, nslabs=1, devices=None, slab_begin=0, slab_count=None,
                 halo_mode=capi.HALO_REFERENCE_LAG2, bc_mode=capi.BC_LITERAL, scale_physical_size=False,
                 kernel_variant=capi.KERNEL_AUTO, batch=0, dims=None, arithmetic=capi.ARITH_KERNEL):
        """dims = (dimx, dimy, dimz): general Cartesian rank grid like init_global_grid's (ranks in MPI Cartesian order,
        z fastest; in-process or one process per GPU). Default: z-slabs (1, 1, nslabs)."""
        self._L = capi.lib()
        self.n = (int(nx), int(ny), int(nz))
        if dims is not None:
            dims = tuple(int(d) for d in dims)
            nslabs = dims[0] * dims[1] * dims[2]
        self.dims = dims if dims is not None else (1, 1, int(nslabs))
        self.nslabs = int(nslabs)
        if dims is not None and slab_count is None and devices is None:
            devices = [0] * self.nslabs
        self.slab_begin = int(slab_begin)
        self.slab_count = self.nslabs if slab_count is None else int(slab_count)
        devices = list(devices) if devices is not None else [0] * self.slab_count
        if len(devices) != self.slab_count:
            raise ValueError("need one device ordinal per hosted slab")
        self._devs = (C.c_int * self.slab_count)(*devices)
        cfg = capi.Diff3DConfig(self.n[0], self.n[1], self.n[2], self.nslabs, self.slab_begin, self.slab_count,
                                C.cast(self._devs, C.POINTER(C.c_int)), halo_mode, bc_mode,
                                int(bool(scale_physical_size)), kernel_variant, batch, self.dims[0], self.dims[1],
                                int(arithmetic))
        self._h = C.c_void_p()
        capi.check(self._L.b2s_diff3d_create(C.byref(self._h), C.byref(cfg)))
        p = capi.Diff3DParams()
        capi.check(self._L.b2s_diff3d_get_params(self._h, C.byref(p)))
        self.params = p
        for k in ("lx", "ly", "lz", "dx", "dy", "dz", "dt", "dtau", "total_N", "nx_g", "ny_g", "nz_g"):
            setattr(self, k, getattr(p, k))

    def close(self):
        if getattr(self, "_h", None):
            self._L.b2s_diff3d_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- set-up --------------------------------------------------------------------------------------------
    def init_gaussian(self):
        capi.check(self._L.b2s_diff3d_init_gaussian(self._h))

    def set_initial(self, Ht):
        a = np.asfortranarray(Ht, dtype=np.float64)
        assert a.size == self.slab_count * self.n[0] * self.n[1] * self.n[2]
        capi.check(self._L.b2s_diff3d_set_initial(self._h, capi.ptr(a)))

    def ipc_export(self):
        n = self._L.b2s_diff3d_ipc_blob_bytes()
        buf = C.create_string_buffer(n)
        capi.check(self._L.b2s_diff3d_ipc_export(self._h, buf))
        return buf.raw

    def ipc_connect(self, blobs):
        raw = b"".join(blobs)
        capi.check(self._L.b2s_diff3d_ipc_connect(self._h, C.c_char_p(raw), len(blobs)))

    # ---- the PT loop ----------------------------------------------------------------------------------------
    def solve_timestep(self, tol, iter_max=100000):
        it, err = C.c_int(), C.c_double()
        capi.check(self._L.b2s_diff3d_solve_timestep(self._h, tol, int(iter_max), C.byref(it), C.byref(err)))
        return it.value, err.value

    def iterate(self, n, want_hist=True):
        hist = np.zeros(n) if want_hist else None
        capi.check(self._L.b2s_diff3d_iterate(self._h, int(n), capi.ptr(hist) if want_hist else None))
        return hist

    def advance_time(self):
        capi.check(self._L.b2s_diff3d_advance_time(self._h))

    def run(self, ttot=1.0, tol=1e-8, iter_max=100000):
        cap = 4096
        its = (C.c_int * cap)()
        ns = C.c_int()
        capi.check(self._L.b2s_diff3d_run(self._h, ttot, tol, int(iter_max), its, cap, C.byref(ns)))
        return list(its[:ns.value])

    # ---- data access ----------------------------------------------------------------------------------------
    def get(self, which, slab=None):
        slab = self.slab_begin if slab is None else slab
        a = np.zeros(self.n, dtype=np.float64, order="F")
        capi.check(self._L.b2s_diff3d_get_field(self._h, slab, {"Ht": 0, "Htau": 1, "Htau2": 2}[which], capi.ptr(a)))
        return a

    def gather(self):
        """gather!(Array(Ht), H_g): (nx*dimx, ny*dimy, nz*dimz) -- for z-slabs (nx, ny, nz*slab_count) --, every hosted
        rank's whole local array at its place in the rank grid."""
        if self.dims[0] * self.dims[1] > 1:
            shape = tuple(n * d for n, d in zip(self.n, self.dims))
        else:
            shape = (self.n[0], self.n[1], self.n[2] * self.slab_count)
        a = np.zeros(shape, dtype=np.float64, order="F")
        capi.check(self._L.b2s_diff3d_gather(self._h, capi.ptr(a)))
        return a

    def device_ptr(self, which, slab=None):
        slab = self.slab_begin if slab is None else slab
        p = C.c_void_p()
        capi.check(self._L.b2s_diff3d_device_ptr(self._h, slab, {"Ht": 0, "Htau": 1, "Htau2": 2}[which], C.byref(p)))
        return p.value

    def upload_state(self, Ht_host, slab=None):
        slab = self.slab_begin if slab is None else slab
        capi.check(self._L.b2s_diff3d_upload_state(self._h, slab, capi.ptr(Ht_host)))

    def download_state(self, out_host, slab=None):
        slab = self.slab_begin if slab is None else slab
        capi.check(self._L.b2s_diff3d_download_state(self._h, slab, capi.ptr(out_host)))

    def download_state_async(self, out_host, slab=None):
        """Pipelined download (separate copy stream): overlaps with the next upload_state; valid after sync()."""
        slab = self.slab_begin if slab is None else slab
        capi.check(self._L.b2s_diff3d_download_state_async(self._h, slab, capi.ptr(out_host)))

    def upload_state_async(self, Ht_host, slab=None):
        """Stage the NEXT job's state on the copy stream (overlaps with running iterations); commit_upload() makes it
        the current state."""
        slab = self.slab_begin if slab is None else slab
        capi.check(self._L.b2s_diff3d_upload_state_async(self._h, slab, capi.ptr(Ht_host)))

    def commit_upload(self, slab=None):
        slab = self.slab_begin if slab is None else slab
        capi.check(self._L.b2s_diff3d_commit_upload(self._h, slab))

    def sync(self):
        capi.check(self._L.b2s_diff3d_sync(self._h))

    def stats(self):
        n, ms = C.c_longlong(), C.c_double()
        capi.check(self._L.b2s_diff3d_stats(self._h, C.byref(n), C.byref(ms)))
        return n.value, ms.value


def diffusion_3D_kernel_programming(*, nx, ny, nz, ttot=1.0, tol=1e-8, use_shared_memory=True, do_vis=False,
                                    verbose=True, init_and_finalize_MPI=True, scale_physical_size=False, nslabs=1,
                                    devices=None, halo_mode=capi.HALO_REFERENCE_LAG2, bc_mode=capi.BC_LITERAL,
                                    kernel_variant=capi.KERNEL_AUTO, return_iters=False, dims=None,
                                    arithmetic=capi.ARITH_KERNEL):
    """Drop-in for scripts-part1/part1_kernel_programming.jl:99.  Returns (X_g, H_g, BenchResults).

    `use_shared_memory` selects the staged (TMA) or the direct kernel; results are bit-identical either way.
    `nslabs`/`devices` replace the MPI rank count: dims = (1, 1, nslabs) z-slabs, one per device entry; `dims` =
    (dimx, dimy, dimz) selects ImplicitGlobalGrid's general rank grid instead (the published 2x2x1 / 2x2x2 layouts).
    """
    if dims is not None:
        nslabs = int(dims[0]) * int(dims[1]) * int(dims[2])
        if devices is None:
            devices = [0] * nslabs
    kv = kernel_variant
    if kv == capi.KERNEL_AUTO and not use_shared_memory:
        kv = capi.KERNEL_DIRECT
    s = Diffusion3D(nx, ny, nz, nslabs=nslabs, devices=devices, halo_mode=halo_mode, bc_mode=bc_mode,
                    scale_physical_size=scale_physical_size, kernel_variant=kv, dims=dims, arithmetic=arithmetic)
    try:
        s.init_gaussian()
        iter_max = 100000  # :130
        stop = ttot - s.dt
        nt = 0 if stop < 0 else int(np.floor(stop / s.dt + 1e-9)) + 1  # length(0:dt:ttot-dt), :166
        iters, timed_iter_total, tic = [], 0, time.time()
        for iter_outer in range(nt):
            if verbose:
                print(f"Iter: {iter_outer}")
            if iter_outer == 3:  # manual warmup, :170-176
                if verbose:
                    print("Starting to measure")
                tic, timed_iter_total = time.time(), 0
            it, err = s.solve_timestep(tol, iter_max)
            if verbose:
                print(f"Converged after {it} iterations." if err <= tol
                      else f"Couldn't converge within {iter_max} iterations.")
            timed_iter_total += it
            iters.append(it)
            s.advance_time()
        s.sync()
        dt_wall = time.time() - tic  # toc() comes before gather! in the reference (:206,223)
        H_g = s.gather()
        cells = (nx - 2) * (ny - 2) * (nz - 2)
        work = float(nslabs) * timed_iter_total * (25 + 2) * cells  # :210
        mem = float(nslabs) * timed_iter_total * ((6 + 1) if use_shared_memory else (14 + 1)) * 8 * cells  # :212-214
        res = BenchResults(dt_wall, work, work / dt_wall if dt_wall > 0 else float("nan"), mem,
                           work / mem if mem else float("nan"), mem / dt_wall if dt_wall > 0 else float("nan"))
        if verbose:
            print(f"Finished after {nt - 2} outer iterations in {dt_wall:3.3f} seconds of compute!")
        X_g = np.linspace(0 + s.dx / 2, s.lx - s.dx / 2, nx * s.dims[0])  # LinRange(dx/2, lx-dx/2, nx*dims[1]), :221
        if return_iters:
            return X_g, H_g, res, iters
        return X_g, H_g, res
    finally:
        s.close()


def diffusion_3D_array_programming(*, nx, ny, nz, do_vis=False, verbose=True, init_and_finalize_MPI=True, nslabs=1,
                                   devices=None, dims=None):
    """Drop-in for scripts-part1/part1_array_programming.jl:20 (ttot = 1, tol = 1e-8 hard-wired there, :23,39): the array
    version's own arithmetic (:9-18: q = D*d(Htau)/dx, divisions instead of reciprocals), Htau updated in place (its frame
    keeps Ht's values) and update_halo!(Htau) after the update (:66-67).  Returns (X_g, H_g)."""
    X_g, H_g, _ = diffusion_3D_kernel_programming(nx=nx, ny=ny, nz=nz, ttot=1.0, tol=1e-8, verbose=verbose, nslabs=nslabs,
                                                  devices=devices, halo_mode=capi.HALO_CONSISTENT, dims=dims,
                                                  arithmetic=capi.ARITH_ARRAY)
    return X_g, H_g


def main(argv):
    """scripts-part1/part1.jl:25-60: [cpu/gpu] [array/kernel] [nx ny nz] [bench]. `cpu` is rejected: no CPU path."""
    args = list(argv)
    if args and args[0] in ("cpu", "gpu"):
        if args[0] == "cpu":
            raise SystemExit("this build has no CPU backend (the reference's Threads path is the oracle/baseline only)")
        args = args[1:]
    version = "kernel"
    if args and args[0] in ("array", "kernel"):
        version, args = args[0], args[1:]
    nx = ny = nz = 32
    if len(args) >= 3 and all(a.isdigit() for a in args[:3]):
        nx, ny, nz = (int(a) for a in args[:3])
        args = args[3:]
    if version == "array":
        diffusion_3D_array_programming(nx=nx, ny=ny, nz=nz)
    else:
        _, _, r = diffusion_3D_kernel_programming(nx=nx, ny=ny, nz=nz)
        print(r)


if __name__ == "__main__":
    import sys
    main(sys.argv[1:])
