# B200Stencil.jl -- ccall binding of libb200stencil.so (include/b200stencil.h) and drop-in replacements for the hot-path
# entry points of ntselepidis/FinalProjectRepo.jl.  NOT executable in the build image (no Julia there): every ccall
# tuple and every struct below is checked mechanically against the header by tests/test_host_cpu.py
# (test_julia_ccall_signatures_match_header); the Python/ctypes mirror (finalprojectrepo.jl_b200/part1.py, part2.py)
# exercises the same symbols in the GPU tests.
#
# Usage inside the reference repository:
#     include("B200Stencil.jl"); using .B200Stencil
#     # scripts-part1/part1.jl:51-52
#     X_g, H_g, bench = B200Stencil.diffusion_3D_kernel_programming(; nx=512, ny=512, nz=512, ttot=1.0, tol=1e-8)
#     # scripts-part2/part2.jl:187 -- u, f are CuArray{Float64,2}
#     r_rms = B200Stencil.MGsolve_2DPoisson!(S, W, h, 0.0, tol, niters, false; prealloc_dict=pre)
#     # scripts-part2/part2.jl:271
#     out = B200Stencil.navier_stokes_2D(; opt=B200Stencil.SimIn_t(), verbose=false)
module B200Stencil

using CUDA

const lib = get(ENV, "B200STENCIL_LIB", joinpath(@__DIR__, "..", "libb200stencil.so"))

struct B2SError <: Exception
    code::Cint
    msg::String
end

last_error() = unsafe_string(ccall((:b2s_last_error, lib), Cstring, ()))
check(rc::Cint) = rc == 0 ? nothing : throw(B2SError(rc, last_error()))

# ---- Part 1 ---------------------------------------------------------------------------------------------------
struct Diff3DConfig            # b2s_diff3d_config
    nx::Cint
    ny::Cint
    nz::Cint
    nslabs_total::Cint
    slab_begin::Cint
    slab_count::Cint
    devices::Ptr{Cint}
    halo_mode::Cint
    bc_mode::Cint
    scale_physical_size::Cint
    kernel_variant::Cint
    batch::Cint
    dimx::Cint                 # general Cartesian decomposition (0/1: z-slabs)
    dimy::Cint
    arithmetic::Cint           # 0 kernel-programming version, 1 array-programming version
end

struct Diff3DParams            # b2s_diff3d_params
    lx::Cdouble
    ly::Cdouble
    lz::Cdouble
    dx::Cdouble
    dy::Cdouble
    dz::Cdouble
    dt::Cdouble
    dtau::Cdouble
    total_N::Cdouble
    nx_g::Cint
    ny_g::Cint
    nz_g::Cint
end

struct BenchResults            # scripts-part1/part1_kernel_programming.jl:22-29
    Δt::Float64
    Work::Float64
    Performance::Float64
    Memory::Float64
    Intensity::Float64
    Throughput::Float64
end

const HALO_REFERENCE_LAG2 = 0
const HALO_CONSISTENT = 1
const KERNEL_AUTO = 0
const KERNEL_DIRECT = 1
const ARITH_KERNEL = 0
const ARITH_ARRAY = 1

# One process per GPU (the `mpiexecjl -np N julia part1.jl` shape): exchange the CUDA-IPC blobs of all ranks' arenas over
# the caller's MPI communicator and map the neighbours. `MPI` is the caller's MPI.jl module (kept out of this module's
# dependencies); the barrier after the connect is required before the first iteration (b200stencil.h).
function connect_ranks!(h::Ptr{Cvoid}, MPI, comm)
    n = ccall((:b2s_diff3d_ipc_blob_bytes, lib), Csize_t, ())
    blob = Vector{UInt8}(undef, n)
    check(ccall((:b2s_diff3d_ipc_export, lib), Cint, (Ptr{Cvoid}, Ptr{UInt8}), h, blob))
    all = MPI.Allgather(blob, comm)
    check(ccall((:b2s_diff3d_ipc_connect, lib), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Cint), h, all, MPI.Comm_size(comm)))
    MPI.Barrier(comm)
    return nothing
end

# Shared body of the two entry points of scripts-part1. With `mpi = (MPI, comm)` this process hosts ONE rank (slab
# `MPI.Comm_rank(comm)` on device `devices[1]`) of an N-rank job; otherwise all ranks are hosted in-process, one per entry
# of `devices`.
function _diffusion_3D(; nx, ny, nz, ttot, tol, kernel_variant, halo_mode, bc_mode, arithmetic, scale_physical_size,
                       devices::Vector{Cint}, dimx::Integer, dimy::Integer, verbose, use_shared_memory, mpi=nothing)
    me = mpi === nothing ? 0 : mpi[1].Comm_rank(mpi[2])
    N = mpi === nothing ? length(devices) : mpi[1].Comm_size(mpi[2])
    hosted = mpi === nothing ? N : 1
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve devices begin
        cfg = Diff3DConfig(nx, ny, nz, N, mpi === nothing ? 0 : me, hosted, pointer(devices), halo_mode, bc_mode,
                           scale_physical_size ? 1 : 0, kernel_variant, 0, dimx, dimy, arithmetic)
        check(ccall((:b2s_diff3d_create, lib), Cint, (Ref{Ptr{Cvoid}}, Ref{Diff3DConfig}), h, cfg))
    end
    try
        p = Ref{Diff3DParams}()
        check(ccall((:b2s_diff3d_get_params, lib), Cint, (Ptr{Cvoid}, Ref{Diff3DParams}), h[], p))
        check(ccall((:b2s_diff3d_init_gaussian, lib), Cint, (Ptr{Cvoid},), h[]))
        mpi === nothing || connect_ranks!(h[], mpi[1], mpi[2])
        dt = p[].dt
        iter_max = 100_000
        iter_outer = 0; timed_iter_total = 0; tic = time()
        for t in 0:dt:ttot-dt
            verbose && me == 0 && println("Iter: $(iter_outer)")
            if iter_outer == 3
                verbose && me == 0 && println("Starting to measure")
                tic = time(); timed_iter_total = 0
            end
            it = Ref{Cint}(0); err = Ref{Cdouble}(0.0)
            check(ccall((:b2s_diff3d_solve_timestep, lib), Cint, (Ptr{Cvoid}, Cdouble, Cint, Ref{Cint}, Ref{Cdouble}),
                        h[], tol, iter_max, it, err))
            if verbose && me == 0
                println(err[] <= tol ? "Converged after $(it[]) iterations." : "Couldn't converge within $iter_max iterations.")
            end
            timed_iter_total += it[]
            iter_outer += 1
            check(ccall((:b2s_diff3d_advance_time, lib), Cint, (Ptr{Cvoid},), h[]))   # Ht .= Hτ
        end
        check(ccall((:b2s_diff3d_sync, lib), Cint, (Ptr{Cvoid},), h[]))
        Δt = time() - tic                                                                   # toc() precedes gather!, :206,223
        dimz = N ÷ (max(dimx, 1) * max(dimy, 1))
        # hosted ranks only: (nx*dimx, ny*dimy, nz*dimz) in-process (:144); one local array with MPI (gather! is the caller's)
        H_g = mpi === nothing ? zeros(nx * max(dimx, 1), ny * max(dimy, 1), nz * dimz) : zeros(nx, ny, nz)
        check(ccall((:b2s_diff3d_gather, lib), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), h[], H_g))
        cells = (nx - 2) * (ny - 2) * (nz - 2)
        Work = N * timed_iter_total * (25 + 2) * cells
        Memory = N * timed_iter_total * ((use_shared_memory ? 6 : 14) + 1) * sizeof(Float64) * cells
        X_g = LinRange(0 + p[].dx / 2, p[].lx - p[].dx / 2, nx * max(dimx, 1))                 # nx * dims[1], :221
        return X_g, H_g, BenchResults(Δt, Work, Work / Δt, Memory, Work / Memory, Memory / Δt)
    finally
        ccall((:b2s_diff3d_destroy, lib), Cint, (Ptr{Cvoid},), h[])
    end
end

"""
Drop-in for `diffusion_3D_kernel_programming` (scripts-part1/part1_kernel_programming.jl:99-228).
`devices` replaces the MPI ranks: one rank per listed CUDA device (ordinals may repeat), driven from this process.
By default the ranks are z-slabs (dims = (1,1,N)); `dimx`, `dimy` select ImplicitGlobalGrid's general decomposition
(dims = (dimx, dimy, N ÷ (dimx*dimy)), ranks in MPI Cartesian order). `mpi = (MPI, comm)` runs one rank per process.
"""
function diffusion_3D_kernel_programming(; nx, ny, nz, ttot=1.0, tol=1e-8, use_shared_memory=true, do_vis=false,
                                         verbose=true, init_and_finalize_MPI=false, scale_physical_size=false,
                                         devices::Vector{Cint}=Cint[0], halo_mode::Integer=HALO_REFERENCE_LAG2,
                                         bc_mode::Integer=0, dimx::Integer=1, dimy::Integer=1, mpi=nothing)
    return _diffusion_3D(; nx=nx, ny=ny, nz=nz, ttot=ttot, tol=tol,
                         kernel_variant=use_shared_memory ? KERNEL_AUTO : KERNEL_DIRECT, halo_mode=halo_mode,
                         bc_mode=bc_mode, arithmetic=ARITH_KERNEL, scale_physical_size=scale_physical_size, devices=devices,
                         dimx=dimx, dimy=dimy, verbose=verbose, use_shared_memory=use_shared_memory, mpi=mpi)
end

"""
Drop-in for `diffusion_3D_array_programming` (scripts-part1/part1_array_programming.jl:20-92): ttot = 1, tol = 1e-8 are
hard-wired there (:23,39); the array version's own arithmetic (:9-18), in-place update of Hτ, `update_halo!(Hτ)` after
the update (:66-67). Returns `(X_g, H_g)`.
"""
function diffusion_3D_array_programming(; nx, ny, nz, do_vis=false, verbose=true, init_and_finalize_MPI=false,
                                        devices::Vector{Cint}=Cint[0], dimx::Integer=1, dimy::Integer=1, mpi=nothing)
    X_g, H_g, _ = _diffusion_3D(; nx=nx, ny=ny, nz=nz, ttot=1.0, tol=1e-8, kernel_variant=KERNEL_DIRECT,
                                halo_mode=HALO_CONSISTENT, bc_mode=0, arithmetic=ARITH_ARRAY, scale_physical_size=false,
                                devices=devices, dimx=dimx, dimy=dimy, verbose=verbose, use_shared_memory=false, mpi=mpi)
    return X_g, H_g
end

"""`main()` of scripts-part1/part1.jl:25-60: `[cpu/gpu] [array/kernel] [nx ny nz] [bench]`. `cpu` is rejected: this
library has no CPU path (the reference's Threads backend stays what it is)."""
function main(args::Vector{String}=ARGS)
    a = copy(args)
    if !isempty(a) && a[1] in ("cpu", "gpu")
        a[1] == "cpu" && error("B200Stencil has no CPU backend: run the reference's Threads path for `cpu`")
        popfirst!(a)
    end
    version = "kernel"
    if !isempty(a) && a[1] in ("array", "kernel")
        version = popfirst!(a)
    end
    nx = ny = nz = 32
    if length(a) >= 3 && all(x -> tryparse(Int, x) !== nothing, a[1:3])
        nx, ny, nz = parse.(Int, a[1:3])
        a = a[4:end]
    end
    bench = !isempty(a) && a[1] == "bench"
    if version == "array"
        return diffusion_3D_array_programming(; nx=nx, ny=ny, nz=nz, verbose=!bench)
    end
    X_g, H_g, r = diffusion_3D_kernel_programming(; nx=nx, ny=ny, nz=nz, verbose=!bench)
    bench && println(r)
    return X_g, H_g, r
end

"""L0: one launch of the fused step kernel on caller-owned CuArrays (replaces the `@parallel diffusion_3D_step_τ…`
call sites, scripts-part1/part1_kernel_programming.jl:181,186)."""
function diffusion_3D_step_τ!(Ht::CuArray{Float64,3}, Hτ::CuArray{Float64,3}, Hτ2::CuArray{Float64,3},
                              dHdτ::CuArray{Float64,3}, dτ, _dt, _dx, _dy, _dz, D_dx, D_dy, D_dz)
    nx, ny, nz = size(Hτ)
    check(ccall((:b2s_diffusion3d_step_tau, lib), Cint,
                (CuPtr{Cdouble}, CuPtr{Cdouble}, CuPtr{Cdouble}, CuPtr{Cdouble}, Cint, Cint, Cint, Cdouble, Cdouble, Cdouble,
                 Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, CuPtr{Cdouble}, Cint, Ptr{Cvoid}),
                Ht, Hτ, Hτ2, dHdτ, nx, ny, nz, dτ, _dt, _dx, _dy, _dz, D_dx, D_dy, D_dz, 1.0, CU_NULL, 0,
                CUDA.stream().handle))
    return nothing
end

# ---- Part 2 ---------------------------------------------------------------------------------------------------
@enum CoarseSolver_t jacobi = 0 conjugate_gradient = 1          # multigrid.jl:10-13
@enum ExecutionPolicy_t serial = 0 parallel = 1 parallel_shmem = 2   # part2_utils.jl:4-8
@enum Smoother_t damped_jacobi = 0 red_black_gauss_seidel = 1   # variant A (reference) / variant B (this library)
@enum Restriction_t injection = 0 full_weighting = 1

mutable struct MGOpt                                            # multigrid.jl:16-22 (+ the variant-B switches)
    coarse_solve_size::Int
    coarse_solver::CoarseSolver_t
    execution_policy::ExecutionPolicy_t
    smoother::Smoother_t
    restriction::Restriction_t
    MGOpt() = new(5, jacobi, parallel_shmem, damped_jacobi, injection)
end

struct MGConfig                                                 # b2s_mg_config
    nx::Cint
    ny::Cint
    coarse_solve_size::Cint
    coarse_solver::Cint
    smoother::Cint
    restriction::Cint
    device::Cint
    use_graph::Cint
    smem_levels::Cint
    fuse_sweeps::Cint
end

mgconfig(nx, ny, opt::MGOpt) = MGConfig(nx, ny, opt.coarse_solve_size, Int(opt.coarse_solver), Int(opt.smoother),
                                        Int(opt.restriction), CUDA.deviceid(CUDA.device()), 1, 1, 1)

mutable struct MGHandle
    ptr::Ptr{Cvoid}
    nx::Int
    ny::Int
end

"""preallocate_buffers(nx, ny) (multigrid.jl:25-38): level table, work arrays and the captured V-cycle graph."""
function preallocate_buffers(nx, ny; opt=MGOpt())
    opt.execution_policy == serial && error("execution policy serial is a CPU-only debug path")   # multigrid.jl:233-236
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:b2s_mg_create, lib), Cint, (Ref{Ptr{Cvoid}}, Ref{MGConfig}), h, mgconfig(nx, ny, opt)))
    hd = MGHandle(h[], nx, ny)
    finalizer(x -> ccall((:b2s_mg_destroy, lib), Cint, (Ptr{Cvoid},), x.ptr), hd)
    return hd
end

"""r_rms = MGsolve_2DPoisson!(u, f, h, c, tol, niters, apply_BCs; opt, verbose, prealloc_dict)  (multigrid.jl:41-84)"""
function MGsolve_2DPoisson!(u::CuArray{Float64,2}, f::CuArray{Float64,2}, h::Float64, c::Float64, tol::Float64,
                            niters::Int, apply_BCs::Bool; opt=MGOpt(), verbose=false, prealloc_dict=nothing)
    nx, ny = size(u)
    hd = prealloc_dict === nothing ? preallocate_buffers(nx, ny; opt=opt) : prealloc_dict
    CUDA.synchronize()
    r = Ref{Cdouble}(0.0); nc = Ref{Cint}(0)
    hist = zeros(max(niters, 1))
    check(ccall((:b2s_mg_solve, lib), Cint,
                (Ptr{Cvoid}, CuPtr{Cdouble}, CuPtr{Cdouble}, Cdouble, Cdouble, Cdouble, Cint, Cint, Ref{Cdouble}, Ref{Cint},
                 Ptr{Cdouble}), hd.ptr, u, f, h, c, tol, niters, apply_BCs ? 1 : 0, r, nc, hist))
    if verbose
        for i in 1:nc[]
            println("Vcycle iter $i: r_rms / f_rms = $(hist[i])")
        end
    end
    if nc[] > 0 && nc[] == niters && !(hist[nc[]] < tol)
        @warn "MGsolve_2DPoisson! did not converge" tol niters       # multigrid.jl:78-80: a warning, not an error
    end
    return r[]
end

"""res_rms = Vcycle_2DPoisson!(u_f, rhs, h, c, tol, coarse_solve_size, coarse_solver, execution_policy, apply_BCs;
prealloc_dict)  (multigrid.jl:91-170): exactly one V-cycle."""
function Vcycle_2DPoisson!(u_f::CuArray{Float64,2}, rhs::CuArray{Float64,2}, h, c, tol, coarse_solve_size,
                           coarse_solver::CoarseSolver_t, execution_policy::ExecutionPolicy_t, apply_BCs::Bool;
                           prealloc_dict=nothing)
    nx, ny = size(u_f)
    ((nx - 1) % 2 != 0 || (ny - 1) % 2 != 0) && error("ERROR:not a power of 2")              # multigrid.jl:95-97
    hd = prealloc_dict
    if hd === nothing
        opt = MGOpt(); opt.coarse_solve_size = coarse_solve_size; opt.coarse_solver = coarse_solver
        opt.execution_policy = execution_policy
        hd = preallocate_buffers(nx, ny; opt=opt)
    end
    CUDA.synchronize()
    r = Ref{Cdouble}(0.0)
    check(ccall((:b2s_mg_vcycle, lib), Cint,
                (Ptr{Cvoid}, CuPtr{Cdouble}, CuPtr{Cdouble}, Cdouble, Cdouble, Cdouble, Cint, Ref{Cdouble}),
                hd.ptr, u_f, rhs, h, c, tol, apply_BCs ? 1 : 0, r))
    return r[]
end

"""MG-preconditioned CG (extension, no counterpart in the reference): (r_rms, iterations). The handle must have been
created with full-weighting restriction (`opt.restriction = full_weighting`)."""
function mg_pcg!(hd::MGHandle, u::CuArray{Float64,2}, f::CuArray{Float64,2}, h::Float64, c::Float64, tol::Float64, maxit::Int;
                 relative_to_rhs::Bool=false)
    CUDA.synchronize()
    r = Ref{Cdouble}(0.0); it = Ref{Cint}(0)
    check(ccall((:b2s_mg_pcg_solve2, lib), Cint,
                (Ptr{Cvoid}, CuPtr{Cdouble}, CuPtr{Cdouble}, Cdouble, Cdouble, Cdouble, Cint, Cint, Ref{Cdouble}, Ref{Cint}),
                hd.ptr, u, f, h, c, tol, maxit, relative_to_rhs ? 1 : 0, r, it))
    return r[], Int(it[])
end

"""res_rms = cg!(x_in, b, hx, hy, c, tol, Nmax; execution_policy, verbose)  (krylov.jl:55-91)"""
function cg!(x_in::CuArray{Float64,2}, b::CuArray{Float64,2}, hx, hy, c, tol, Nmax; execution_policy=parallel_shmem,
             verbose=false)
    nx, ny = size(b)
    r = Ref{Cdouble}(0.0); it = Ref{Cint}(0)
    check(ccall((:b2s_cg_solve, lib), Cint,
                (CuPtr{Cdouble}, CuPtr{Cdouble}, Cdouble, Cdouble, Cdouble, Cdouble, Cint, Cint, Cint, Cint, Ref{Cdouble},
                 Ref{Cint}, Ptr{Cvoid}), x_in, b, hx, hy, c, tol, Nmax, nx, ny, Int(execution_policy), r, it,
                CUDA.stream().handle))
    return r[]
end

"""r_rms = iteration_2DPoisson!(u, f, h, c, res, execution_policy; alpha)  (multigrid.jl:245-258)"""
function iteration_2DPoisson!(u::CuArray{Float64,2}, f, h, c, res, execution_policy; alpha=4.0 / 5.0)
    nx, ny = size(u)
    r = Ref{Cdouble}(0.0)
    check(ccall((:b2s_iteration2d, lib), Cint,
                (CuPtr{Cdouble}, CuPtr{Cdouble}, Cdouble, Cdouble, CuPtr{Cdouble}, Cint, Cint, Cdouble, Cint, Ref{Cdouble},
                 Ptr{Cvoid}), u, f, h, c, res, nx, ny, alpha, Int(execution_policy), r, CUDA.stream().handle))
    return r[]
end

function residual_2DPoisson_wrapper!(u::CuArray{Float64,2}, f, h, c, res, execution_policy)   # multigrid.jl:223-238
    nx, ny = size(u)
    check(ccall((:b2s_residual2d, lib), Cint,
                (CuPtr{Cdouble}, CuPtr{Cdouble}, Cdouble, Cdouble, CuPtr{Cdouble}, Cint, Cint, Cint, Ptr{Cvoid}),
                u, f, h, c, res, nx, ny, Int(execution_policy), CUDA.stream().handle))
end

function restrict_wrapper!(fine::CuArray{Float64,2}, coarse, apply_BCs, execution_policy)     # multigrid.jl:344-358
    nx, ny = size(fine)
    check(ccall((:b2s_restrict_inject2d, lib), Cint, (CuPtr{Cdouble}, CuPtr{Cdouble}, Cint, Cint, Cint, Ptr{Cvoid}),
                fine, coarse, nx, ny, apply_BCs ? 1 : 0, CUDA.stream().handle))
end

function prolongate_wrapper!(coarse::CuArray{Float64,2}, fine, apply_BCs, execution_policy)   # multigrid.jl:451-472
    nx, ny = size(fine)
    check(ccall((:b2s_prolongate2d, lib), Cint, (CuPtr{Cdouble}, CuPtr{Cdouble}, Cint, Cint, Cint, Ptr{Cvoid}),
                coarse, fine, nx, ny, apply_BCs ? 1 : 0, CUDA.stream().handle))
end

function matrix_free_matvec_prod_wrapper!(p::CuArray{Float64,2}, hx, hy, c, p_hat, execution_policy)   # krylov.jl:37-52
    nx, ny = size(p)
    check(ccall((:b2s_matvec2d, lib), Cint,
                (CuPtr{Cdouble}, Cdouble, Cdouble, Cdouble, CuPtr{Cdouble}, Cint, Cint, Cint, Ptr{Cvoid}),
                p, hx, hy, c, p_hat, nx, ny, Int(execution_policy), CUDA.stream().handle))
    CUDA.synchronize()
end

apply_boundary_conditions!(T::CuArray{Float64,2}) =                                              # part2_utils.jl:21-24
    check(ccall((:b2s_apply_bc2d, lib), Cint, (CuPtr{Cdouble}, Cint, Cint, Cint, Ptr{Cvoid}), T, size(T, 1), size(T, 2), 0,
                CUDA.stream().handle))

# ---- Navier-Stokes driver (scripts-part2/part2.jl) ---------------------------------------------------------------
@enum Init_t cosine random W_from_file                          # part2.jl:24-28

mutable struct SimIn_t                                          # part2.jl:30-46
    k::Float64
    Ra::Float64
    Pr::Float64
    nx::Int
    ny::Int
    ttot::Float64
    beta::Float64
    niters::Int
    tol::Float64
    a_dif::Float64
    a_adv::Float64
    T_init_strategy::Init_t
    W_init_strategy::Init_t
    SimIn_t() = new(1.0, 1.0e6, 1.0e-3, 257, 65, 0.1, 0.0, 50, 1.0e-3, 0.15, 0.4, cosine, random)
end

struct SimOut_t                                                 # part2.jl:49-55
    T::Matrix{Float64}
    W::Matrix{Float64}
    S::Matrix{Float64}
    t_elapsed::Float64
    timed_iters::Float64
end

struct NS2DParams                                               # b2s_ns2d_params
    k::Cdouble
    Ra::Cdouble
    Pr::Cdouble
    nx::Cint
    ny::Cint
    ttot::Cdouble
    beta::Cdouble
    niters::Cint
    tol::Cdouble
    a_dif::Cdouble
    a_adv::Cdouble
end

struct NS2DStepInfo                                             # b2s_ns2d_stepinfo
    dt::Cdouble
    cycles_S::Cint
    cycles_T::Cint
    cycles_W::Cint
    r_S::Cdouble
    r_T::Cdouble
    r_W::Cdouble
end

const NS_FIELD_T = 0
const NS_FIELD_W = 1
const NS_FIELD_S = 2

function _ns_init!(h::Ptr{Cvoid}, which::Integer, scheme::Init_t, nx, ny, W_init)
    if scheme == cosine
        check(ccall((:b2s_ns2d_init_cosine, lib), Cint, (Ptr{Cvoid}, Cint), h, which))
    else
        M = scheme == random ? rand(nx, ny) : (W_init === nothing ? error("W_from_file needs W_init") : Matrix{Float64}(W_init))
        check(ccall((:b2s_ns2d_set_field, lib), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}), h, which, M))
    end
end

"""
Drop-in for `navier_stokes_2D(; opt, verbose, do_vis, testmode)` (scripts-part2/part2.jl:140-262): the whole time loop
(three multigrid solves per step, the fused velocity / stencil-term kernels around them) runs inside the library on the
current CUDA device; returns `SimOut_t` with host matrices like the reference. `W_init` supplies the field for
`W_init_strategy = W_from_file` (the reference reads `test/reftest-files/fortran/Winit.bin`, part2.jl:67-73);
`mgopt` selects the multigrid variant (default: the reference's damped Jacobi + injection); `mg_pcg = true` solves the two
Dirichlet systems of a step (S, W) with MG-preconditioned CG instead of plain V-cycle iteration (needs
`mgopt.restriction = full_weighting`).
"""
function navier_stokes_2D(; opt::SimIn_t=SimIn_t(), verbose=true, do_vis=false, testmode=false, W_init=nothing, mgopt=MGOpt(),
                          mg_pcg::Bool=false)
    nx, ny = opt.nx, opt.ny
    prm = NS2DParams(opt.k, opt.Ra, opt.Pr, nx, ny, opt.ttot, opt.beta, opt.niters, opt.tol, opt.a_dif, opt.a_adv)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:b2s_ns2d_create, lib), Cint, (Ref{Ptr{Cvoid}}, Ref{NS2DParams}, Ref{MGConfig}), h, prm, mgconfig(nx, ny, mgopt)))
    try
        mg_pcg && check(ccall((:b2s_ns2d_set_solver, lib), Cint, (Ptr{Cvoid}, Cint), h[], 1))
        _ns_init!(h[], NS_FIELD_T, opt.T_init_strategy, nx, ny, nothing)
        _ns_init!(h[], NS_FIELD_W, opt.W_init_strategy, nx, ny, W_init)
        tic = 0.0; sim_time = 0.0; step = 0
        while sim_time < opt.ttot
            step == 3 && (tic = time())
            info = Ref{NS2DStepInfo}()
            check(ccall((:b2s_ns2d_step, lib), Cint, (Ptr{Cvoid}, Ref{NS2DStepInfo}), h[], info))
            sim_time += info[].dt
            step += 1
            ((step - 1) % 20 == 0) && verbose && println("time, step: $(sim_time) $(step)")
            testmode && break
        end
        t_elapsed = time() - tic                    # b2s_ns2d_step returns dt to the host, i.e. it has synchronised
        T = zeros(nx, ny); W = zeros(nx, ny); S = zeros(nx, ny)
        check(ccall((:b2s_ns2d_get_field, lib), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}), h[], NS_FIELD_T, T))
        check(ccall((:b2s_ns2d_get_field, lib), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}), h[], NS_FIELD_W, W))
        check(ccall((:b2s_ns2d_get_field, lib), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}), h[], NS_FIELD_S, S))
        println("time, step: $(sim_time) $(step)")
        return SimOut_t(T, W, S, t_elapsed, step - 3)
    finally
        ccall((:b2s_ns2d_destroy, lib), Cint, (Ptr{Cvoid},), h[])
    end
end

end # module
