# B200Stencil.jl -- ccall binding of libb200stencil.so (include/b200stencil.h) and drop-in replacements for the hot-path
# entry points of ntselepidis/FinalProjectRepo.jl.  NOT executable in the build image (no Julia there): syntax-reviewed
# only; the Python/ctypes mirror (finalprojectrepo.jl_b200/part1.py, part2.py) exercises the same symbols in the tests.
#
# Usage inside the reference repository:
#     include("B200Stencil.jl"); using .B200Stencil
#     # scripts-part1/part1.jl:51-52
#     X_g, H_g, bench = B200Stencil.diffusion_3D_kernel_programming(; nx=512, ny=512, nz=512, ttot=1.0, tol=1e-8)
#     # scripts-part2/part2.jl:187 -- u, f are CuArray{Float64,2}
#     r_rms = B200Stencil.MGsolve_2DPoisson!(S, W, h, 0.0, tol, niters, false; prealloc_dict=pre)
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
    nx::Cint; ny::Cint; nz::Cint
    nslabs_total::Cint; slab_begin::Cint; slab_count::Cint
    devices::Ptr{Cint}
    halo_mode::Cint; bc_mode::Cint; scale_physical_size::Cint; kernel_variant::Cint; batch::Cint
    dimx::Cint; dimy::Cint     # general Cartesian decomposition (0/1: z-slabs)
end

struct Diff3DParams            # b2s_diff3d_params
    lx::Cdouble; ly::Cdouble; lz::Cdouble; dx::Cdouble; dy::Cdouble; dz::Cdouble; dt::Cdouble; dtau::Cdouble
    total_N::Cdouble
    nx_g::Cint; ny_g::Cint; nz_g::Cint
end

struct BenchResults            # scripts-part1/part1_kernel_programming.jl:22-29
    Δt::Float64
    Work::Float64
    Performance::Float64
    Memory::Float64
    Intensity::Float64
    Throughput::Float64
end

"""
Drop-in for `diffusion_3D_kernel_programming` (scripts-part1/part1_kernel_programming.jl:99-228).
`devices` replaces the MPI ranks: one rank per listed CUDA device (ordinals may repeat), driven from this process.
By default the ranks are z-slabs (dims = (1,1,N)); `dimx`, `dimy` select ImplicitGlobalGrid's general decomposition
(dims = (dimx, dimy, N ÷ (dimx*dimy)), ranks in MPI Cartesian order).
"""
function diffusion_3D_kernel_programming(; nx, ny, nz, ttot=1.0, tol=1e-8, use_shared_memory=true, do_vis=false,
                                         verbose=true, init_and_finalize_MPI=false, scale_physical_size=false,
                                         devices::Vector{Cint}=Cint[0], halo_mode::Integer=0, bc_mode::Integer=0,
                                         dimx::Integer=1, dimy::Integer=1)
    N = length(devices)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve devices begin
        cfg = Diff3DConfig(nx, ny, nz, N, 0, N, pointer(devices), halo_mode, bc_mode, scale_physical_size ? 1 : 0,
                           use_shared_memory ? 0 : 1, 0, dimx, dimy)
        check(ccall((:b2s_diff3d_create, lib), Cint, (Ref{Ptr{Cvoid}}, Ref{Diff3DConfig}), h, cfg))
    end
    try
        p = Ref{Diff3DParams}()
        check(ccall((:b2s_diff3d_get_params, lib), Cint, (Ptr{Cvoid}, Ref{Diff3DParams}), h[], p))
        check(ccall((:b2s_diff3d_init_gaussian, lib), Cint, (Ptr{Cvoid},), h[]))
        dt = p[].dt
        iter_max = 100_000
        iter_outer = 0; timed_iter_total = 0; tic = time()
        for t in 0:dt:ttot-dt
            verbose && println("Iter: $(iter_outer)")
            if iter_outer == 3
                verbose && println("Starting to measure")
                tic = time(); timed_iter_total = 0
            end
            it = Ref{Cint}(0); err = Ref{Cdouble}(0.0)
            check(ccall((:b2s_diff3d_solve_timestep, lib), Cint, (Ptr{Cvoid}, Cdouble, Cint, Ref{Cint}, Ref{Cdouble}),
                        h[], tol, iter_max, it, err))
            if verbose
                println(err[] <= tol ? "Converged after $(it[]) iterations." : "Couldn't converge within $iter_max iterations.")
            end
            timed_iter_total += it[]
            iter_outer += 1
            check(ccall((:b2s_diff3d_advance_time, lib), Cint, (Ptr{Cvoid},), h[]))   # Ht .= Hτ
        end
        dimz = N ÷ (dimx * dimy)
        H_g = zeros(nx * dimx, ny * dimy, nz * dimz)                                        # zeros(nx*dims[1], …) :144
        check(ccall((:b2s_diff3d_gather, lib), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), h[], H_g))   # also synchronises
        Δt = time() - tic
        cells = (nx - 2) * (ny - 2) * (nz - 2)
        Work = N * timed_iter_total * (25 + 2) * cells
        Memory = N * timed_iter_total * ((use_shared_memory ? 6 : 14) + 1) * sizeof(Float64) * cells
        X_g = LinRange(0 + p[].dx / 2, p[].lx - p[].dx / 2, nx)
        return X_g, H_g, BenchResults(Δt, Work, Work / Δt, Memory, Work / Memory, Memory / Δt)
    finally
        ccall((:b2s_diff3d_destroy, lib), Cint, (Ptr{Cvoid},), h[])
    end
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

mutable struct MGOpt                                            # multigrid.jl:16-22
    coarse_solve_size::Int
    coarse_solver::CoarseSolver_t
    execution_policy::ExecutionPolicy_t
    MGOpt() = new(5, jacobi, parallel_shmem)
end

struct MGConfig                                                 # b2s_mg_config
    nx::Cint; ny::Cint; coarse_solve_size::Cint; coarse_solver::Cint; smoother::Cint; restriction::Cint
    device::Cint; use_graph::Cint; smem_levels::Cint; fuse_sweeps::Cint
end

mutable struct MGHandle
    ptr::Ptr{Cvoid}
    nx::Int
    ny::Int
end

"""preallocate_buffers(nx, ny) (multigrid.jl:25-38): level table, work arrays and the captured V-cycle graph."""
function preallocate_buffers(nx, ny; opt=MGOpt())
    h = Ref{Ptr{Cvoid}}(C_NULL)
    cfg = MGConfig(nx, ny, opt.coarse_solve_size, Int(opt.coarse_solver), 0, 0, CUDA.deviceid(CUDA.device()), 1, 1, 1)
    check(ccall((:b2s_mg_create, lib), Cint, (Ref{Ptr{Cvoid}}, Ref{MGConfig}), h, cfg))
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
    if nc[] == niters && !(hist[nc[]] < tol)
        @warn "MGsolve_2DPoisson! did not converge" tol niters       # multigrid.jl:78-80: a warning, not an error
    end
    return r[]
end

"""MG-preconditioned CG (extension, no counterpart in the reference): (r_rms, iterations). The handle must have been
created with full-weighting restriction (MGConfig.restriction = 1)."""
function mg_pcg!(hd::MGHandle, u::CuArray{Float64,2}, f::CuArray{Float64,2}, h::Float64, c::Float64, tol::Float64, maxit::Int)
    CUDA.synchronize()
    r = Ref{Cdouble}(0.0); it = Ref{Cint}(0)
    check(ccall((:b2s_mg_pcg_solve, lib), Cint,
                (Ptr{Cvoid}, CuPtr{Cdouble}, CuPtr{Cdouble}, Cdouble, Cdouble, Cdouble, Cint, Ref{Cdouble}, Ref{Cint}),
                hd.ptr, u, f, h, c, tol, maxit, r, it))
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

end # module
