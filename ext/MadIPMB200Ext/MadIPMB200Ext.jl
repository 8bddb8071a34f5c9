# MadIPMB200Ext -- Julia glue between MadIPM.jl / MadNLP.jl and libmadipm_b200.so.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia runtime (SURVEY.md fact 3).
# The same C ABI is exercised from Python (madipm_jl_b200/_lib.py, tests/test_gpu_parity.py);
# this file is the binding a MadIPM maintainer would add under ext/ next to ext/MadIPMCUDAExt
# (see INTEGRATION.md). It mirrors the dispatch points of ext/MadIPMCUDAExt/cuda_wrapper.jl:
#
#   reference hook (file:line)                                   -> C ABI entry
#   MadIPM.coo_to_csr            cuda_wrapper.jl:96-106           -> mipm_coo_to_csr
#   MadIPM.build_normal_system   cuda_wrapper.jl:214-234          -> mipm_normal_symbolic
#   MadNLP.compress_jacobian!    cuda_wrapper.jl:32-41            -> gather + mipm_normal_set_jacobian
#   MadIPM.assemble_normal_system! cuda_wrapper.jl:141-156        -> mipm_normal_assemble
#   MadNLP.transfer!             cuda_wrapper.jl:12-24            -> mipm_k2_transfer
#   linear_solver(aug_com; opt)  normalkkt.jl:113-115             -> B200Solver (mipm_ls_analyze)
#   MadNLP.factorize!/solve!     linear_solver.jl:10, normalkkt.jl:210 -> mipm_ls_factorize_async / mipm_ls_solve
#   MadIPM.is_factorized         src/utils.jl:54-62               -> mipm_ls_status
#   src/kernels.jl vector functions on MPCSolver{T,<:CuVector}    -> mipm_set_* / mipm_get_* (fused)
module MadIPMB200Ext

using LinearAlgebra
using CUDA
using CUDA.CUSPARSE
import MadNLP
import MadIPM

const libmadipm = get(ENV, "MADIPM_B200_LIB", "libmadipm_b200.so")

const MIPM_OK = Cint(0)
const MIPM_ERR_NOT_FACTORIZED = Cint(6)
const MIPM_CHOLESKY = Cint(0)
const MIPM_LDL = Cint(1)

mutable struct Handle
    ptr::Ptr{Cvoid}
    function Handle(dev::Integer = CUDA.deviceid(CUDA.device()), stream = CUDA.stream())
        out = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:mipm_create, libmadipm), Cint, (Ref{Ptr{Cvoid}}, Cint, Ptr{Cvoid}),
                   out, dev, reinterpret(Ptr{Cvoid}, stream.handle))
        rc == MIPM_OK || error("mipm_create failed with code $rc (no CPU fallback exists)")
        h = new(out[])
        finalizer(x -> ccall((:mipm_destroy, libmadipm), Cint, (Ptr{Cvoid},), x.ptr), h)
        return h
    end
end

function check(h::Handle, rc::Cint)
    rc == MIPM_OK && return
    msg = unsafe_string(ccall((:mipm_last_error, libmadipm), Cstring, (Ptr{Cvoid},), h.ptr))
    throw(MadNLP.LinearSolverException())  # message: msg (SURVEY 8b: non-zero -> LinearSolverException)
end

devptr(x::CuArray) = reinterpret(Ptr{Cvoid}, pointer(x))

# ---------------------------------------------------------------------------------------
# Linear solver: replaces MadNLPGPU.CUDSSSolver behind MadNLP.AbstractLinearSolver
# ---------------------------------------------------------------------------------------
@kwdef mutable struct B200Options <: MadNLP.AbstractOptions
    b200_algorithm::MadNLP.LinearFactorization = MadNLP.LDL   # MadNLP.CHOLESKY for NormalKKTSystem
    b200_ordering::Int = 0                                     # 0 = nested dissection, 1 = natural
    b200_ir_steps::Int = 0
end

mutable struct B200Solver{T} <: MadNLP.AbstractLinearSolver{T}
    handle::Handle
    tril::CUSPARSE.CuSparseMatrixCSC{T,Int32}   # aug_com: lower CSC, pattern fixed, values rewritten in place
    opt::B200Options
    logger::MadNLP.MadNLPLogger
end

function B200Solver(csc::CUSPARSE.CuSparseMatrixCSC{T,Int32};
                    opt = B200Options(), logger = MadNLP.MadNLPLogger()) where {T}
    h = Handle()
    n = size(csc, 1)
    colptr = Vector(csc.colPtr)      # host copies for the one-time analysis, 1-based Int32
    rowval = Vector(csc.rowVal)
    kind = opt.b200_algorithm == MadNLP.CHOLESKY ? MIPM_CHOLESKY : MIPM_LDL
    check(h, ccall((:mipm_ls_analyze, libmadipm), Cint,
                   (Ptr{Cvoid}, Int64, Ptr{Int32}, Ptr{Int32}, Cint, Cint, Cint, Ptr{Int32}),
                   h.ptr, n, colptr, rowval, 1, kind, opt.b200_ordering, C_NULL))
    return B200Solver{T}(h, csc, opt, logger)
end

function MadNLP.factorize!(M::B200Solver)
    check(M.handle, ccall((:mipm_ls_factorize_async, libmadipm), Cint, (Ptr{Cvoid}, Ptr{Cvoid}),
                          M.handle.ptr, devptr(M.tril.nzVal)))
    return M
end

function MadNLP.solve!(M::B200Solver{T}, x::CuVector{T}) where {T}
    check(M.handle, ccall((:mipm_ls_solve, libmadipm), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint),
                          M.handle.ptr, devptr(x), M.opt.b200_ir_steps))
    return x
end

function MadIPM.is_factorized(M::B200Solver)
    st = Ref{Cint}(0)
    check(M.handle, ccall((:mipm_ls_status, libmadipm), Cint, (Ptr{Cvoid}, Ref{Cint}), M.handle.ptr, st))
    return st[] == MIPM_OK
end

MadNLP.is_inertia(::B200Solver) = true
function MadNLP.inertia(M::B200Solver)
    p, z, n = Ref{Int64}(0), Ref{Int64}(0), Ref{Int64}(0)
    check(M.handle, ccall((:mipm_ls_inertia, libmadipm), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}, Ref{Int64}),
                          M.handle.ptr, p, z, n))
    return (p[], z[], n[])
end
MadNLP.improve!(::B200Solver) = false
MadNLP.introduce(::B200Solver) = "madipm_b200 (supernodal Cholesky / LDL' on B200)"
MadNLP.input_type(::Type{<:B200Solver}) = :csc
MadNLP.default_options(::Type{<:B200Solver}) = B200Options()
MadNLP.is_supported(::Type{<:B200Solver}, ::Type{Float64}) = true
MadNLP.is_supported(::Type{<:B200Solver}, ::Type{T}) where {T} = false

# ---------------------------------------------------------------------------------------
# KKT assembly overrides (same dispatch points as ext/MadIPMCUDAExt/cuda_wrapper.jl)
# ---------------------------------------------------------------------------------------
# One handle per KKT system, created lazily and cached on the object id.
const KKT_HANDLES = IdDict{Any,Handle}()
kkt_handle(kkt) = get!(() -> Handle(), KKT_HANDLES, kkt)

# Symbolic tril(A A'): host arrays in, host arrays out (normalkkt.jl:104 calls it with Vectors).
function MadIPM.build_normal_system(n_rows, n_cols, Jtp::Vector{Int32}, Jtj::Vector{Int32}; handle::Handle = Handle())
    Cp, Cj, nnzC = Ref{Ptr{Int32}}(C_NULL), Ref{Ptr{Int32}}(C_NULL), Ref{Int64}(0)
    check(handle, ccall((:mipm_normal_symbolic, libmadipm), Cint,
                        (Ptr{Cvoid}, Int64, Int64, Ptr{Int32}, Ptr{Int32}, Cint, Ref{Ptr{Int32}}, Ref{Ptr{Int32}}, Ref{Int64}),
                        handle.ptr, n_rows, n_cols, Jtp, Jtj, 1, Cp, Cj, nnzC))
    cp = copy(unsafe_wrap(Array, Cp[], n_rows + 1)); ccall((:mipm_free, libmadipm), Cvoid, (Ptr{Cvoid},), Cp[])
    cj = copy(unsafe_wrap(Array, Cj[], nnzC[]));     ccall((:mipm_free, libmadipm), Cvoid, (Ptr{Cvoid},), Cj[])
    return (cp, cj)
end

function MadNLP.compress_jacobian!(kkt::MadIPM.NormalKKTSystem{T,VT,MT}) where {T,VT,MT<:CUSPARSE.CuSparseMatrixCSC{T,Int32}}
    n_slack = length(kkt.ind_ineq)
    kkt.A.V[end-n_slack+1:end] .= -one(T)
    kkt.AT.nzVal .= kkt.A.V[kkt.A_csr_map]          # one gather, once per solve (solver.jl:167)
    h = kkt_handle(kkt)
    check(h, ccall((:mipm_normal_set_jacobian, libmadipm), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), h.ptr, devptr(kkt.AT.nzVal)))
    return
end

# build_kkt!(::NormalKKTSystem) calls this with D = 1 ./ pr_diag (normalkkt.jl:191-192); the library
# takes pr_diag itself and forms D in the same launch sequence, so we pass kkt.pr_diag through Dx's owner.
function MadNLP.build_kkt!(kkt::MadIPM.NormalKKTSystem{T,VT,MT}) where {T,VT,MT<:CUSPARSE.CuSparseMatrixCSC{T,Int32}}
    h = kkt_handle(kkt)
    check(h, ccall((:mipm_normal_assemble, libmadipm), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint),
                   h.ptr, devptr(kkt.pr_diag), devptr(kkt.aug_com.nzVal), 0))
    return
end

# K2: MadNLP.transfer!(aug_com, aug_raw, aug_csc_map) as a deterministic gather (cuda_wrapper.jl:12-24 races).
# The (I, J) -> CSC map is rebuilt once per KKT system through mipm_k2_symbolic (same pattern as MadNLP's coo_to_csc).
function MadNLP.transfer!(dest::CUSPARSE.CuSparseMatrixCSC{Tv}, src::MadNLP.SparseMatrixCOO{Tv}, map::CuVector{Int}) where {Tv}
    h = get!(KKT_HANDLES, dest) do
        hh = Handle()
        colptr, rowval, m, nnz = Ref{Ptr{Int32}}(C_NULL), Ref{Ptr{Int32}}(C_NULL), Ref{Ptr{Int64}}(C_NULL), Ref{Int64}(0)
        I, J = Vector{Int32}(src.I), Vector{Int32}(src.J)
        check(hh, ccall((:mipm_k2_symbolic, libmadipm), Cint,
                        (Ptr{Cvoid}, Int64, Int64, Ptr{Int32}, Ptr{Int32}, Cint, Ref{Ptr{Int32}}, Ref{Ptr{Int32}}, Ref{Ptr{Int64}}, Ref{Int64}),
                        hh.ptr, size(dest, 1), length(I), I, J, 1, colptr, rowval, m, nnz))
        for p in (colptr[], rowval[], m[]); ccall((:mipm_free, libmadipm), Cvoid, (Ptr{Cvoid},), p); end
        hh
    end
    check(h, ccall((:mipm_k2_transfer, libmadipm), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                   h.ptr, devptr(src.V), devptr(nonzeros(dest))))
    return
end

# ---------------------------------------------------------------------------------------
# Fused vector kernels: specialise src/kernels.jl on GPU solvers
# ---------------------------------------------------------------------------------------
const GPUSolver = MadIPM.MPCSolver{T,VT} where {T,VT<:CuVector{T}}

# C mirror of mipm_mpc_vectors (include/madipm_b200.h); field order must match.
struct MpcVectors
    n::Int64; m::Int64; nlb::Int64; nub::Int64
    index_base::Cint
    ind_lb::Ptr{Cvoid}; ind_ub::Ptr{Cvoid}
    x::Ptr{Cvoid}; xl::Ptr{Cvoid}; xu::Ptr{Cvoid}; zl::Ptr{Cvoid}; zu::Ptr{Cvoid}; f::Ptr{Cvoid}
    y::Ptr{Cvoid}; c::Ptr{Cvoid}; rhs::Ptr{Cvoid}
    jacl::Ptr{Cvoid}
    d::Ptr{Cvoid}; p::Ptr{Cvoid}; w::Ptr{Cvoid}
    corr_lb::Ptr{Cvoid}; corr_ub::Ptr{Cvoid}
    reg::Ptr{Cvoid}; pr_diag::Ptr{Cvoid}; du_diag::Ptr{Cvoid}
    l_diag::Ptr{Cvoid}; u_diag::Ptr{Cvoid}; l_lower::Ptr{Cvoid}; u_lower::Ptr{Cvoid}
end

const SOLVER_HANDLES = IdDict{Any,Handle}()
function solver_handle(s::GPUSolver)
    get!(SOLVER_HANDLES, s) do
        h = kkt_handle(s.kkt)
        k = s.kkt
        v = MpcVectors(s.n, s.m, s.nlb, s.nub, 1, devptr(s.ind_lb), devptr(s.ind_ub),
                       devptr(MadNLP.full(s.x)), devptr(MadNLP.full(s.xl)), devptr(MadNLP.full(s.xu)),
                       devptr(MadNLP.full(s.zl)), devptr(MadNLP.full(s.zu)), devptr(MadNLP.full(s.f)),
                       devptr(s.y), devptr(s.c), devptr(s.rhs), devptr(s.jacl),
                       devptr(MadNLP.full(s.d)), devptr(MadNLP.full(s.p)), devptr(MadNLP.full(s._w1)),
                       devptr(s.correction_lb), devptr(s.correction_ub),
                       devptr(k.reg), devptr(k.pr_diag), devptr(k.du_diag),
                       devptr(k.l_diag), devptr(k.u_diag), devptr(k.l_lower), devptr(k.u_lower))
        check(h, ccall((:mipm_mpc_bind, libmadipm), Cint, (Ptr{Cvoid}, Ref{MpcVectors}), h.ptr, Ref(v)))
        h
    end
end

function MadIPM.set_aug_diagonal_reg!(kkt::MadNLP.AbstractKKTSystem{T}, s::GPUSolver{T}) where {T}
    h = solver_handle(s)
    check(h, ccall((:mipm_set_aug_diagonal_reg, libmadipm), Cint, (Ptr{Cvoid}, Cdouble, Cdouble), h.ptr, s.del_w, s.del_c))
end
function MadIPM.set_predictive_rhs!(s::GPUSolver, kkt::MadNLP.AbstractKKTSystem)
    h = solver_handle(s)
    check(h, ccall((:mipm_set_predictive_rhs, libmadipm), Cint, (Ptr{Cvoid},), h.ptr))
end
function MadIPM.set_correction_rhs!(s::GPUSolver, kkt::MadNLP.AbstractKKTSystem, mu::Float64, clb, cub, ind_lb, ind_ub)
    h = solver_handle(s)
    check(h, ccall((:mipm_set_correction_rhs, libmadipm), Cint, (Ptr{Cvoid}, Cdouble), h.ptr, mu))
end
function MadIPM.get_correction!(s::GPUSolver, clb, cub)
    h = solver_handle(s)
    check(h, ccall((:mipm_get_correction, libmadipm), Cint, (Ptr{Cvoid},), h.ptr))
end
function MadIPM.set_extra_correction!(s::GPUSolver, clb, cub, alpha_p, alpha_d, bmin, bmax, mu)
    h = solver_handle(s)
    check(h, ccall((:mipm_set_extra_correction, libmadipm), Cint,
                   (Ptr{Cvoid}, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble), h.ptr, alpha_p, alpha_d, bmin, bmax, mu))
end
function MadIPM.get_complementarity_measure(s::GPUSolver)
    h = solver_handle(s); out = Ref{Cdouble}(0)
    check(h, ccall((:mipm_get_complementarity_measure, libmadipm), Cint, (Ptr{Cvoid}, Ref{Cdouble}), h.ptr, out))
    return out[]
end
function MadIPM.get_affine_complementarity_measure(s::GPUSolver, alpha_p, alpha_d)
    h = solver_handle(s); out = Ref{Cdouble}(0)
    check(h, ccall((:mipm_get_affine_complementarity_measure, libmadipm), Cint,
                   (Ptr{Cvoid}, Cdouble, Cdouble, Ref{Cdouble}), h.ptr, alpha_p, alpha_d, out))
    return out[]
end
function MadIPM.get_fraction_to_boundary_step(s::GPUSolver, tau)
    h = solver_handle(s); a = zeros(Cdouble, 4); i = zeros(Int64, 4)
    check(h, ccall((:mipm_get_alpha_max, libmadipm), Cint, (Ptr{Cvoid}, Cdouble, Ptr{Cdouble}, Ptr{Int64}), h.ptr, tau, a, i))
    return min(a[1], a[2]), min(a[3], a[4])
end
function MadIPM.apply_step!(s::GPUSolver)
    h = solver_handle(s)
    check(h, ccall((:mipm_apply_step, libmadipm), Cint, (Ptr{Cvoid}, Cdouble, Cdouble, Cdouble), h.ptr, s.alpha_p, s.alpha_d, s.mu))
    s.cnt.k += 1
    return
end
function MadIPM.dual_objective(s::GPUSolver)
    h = solver_handle(s); out = zeros(Cdouble, 5)
    check(h, ccall((:mipm_termination_measures, libmadipm), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), h.ptr, out))
    return out[1]
end

end # module
