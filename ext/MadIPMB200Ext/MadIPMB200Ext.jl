# MadIPMB200Ext -- Julia glue between MadIPM.jl / MadNLP.jl and libmadipm_b200.so.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia runtime (SURVEY.md fact 3). The same C ABI, with
# the same calling convention (one handle per KKT system, index_base = 1, Int32 patterns, Int64 value maps and bound
# indices, library-owned host arrays released with mipm_free), is exercised from Python by
# tests/test_gpu_parity.py::test_julia_calling_convention (madipm_jl_b200/solver.py with index_base = 1).
# This file is the binding a MadIPM maintainer would add under ext/ next to ext/MadIPMCUDAExt (see INTEGRATION.md).
#
# With this extension loaded, no CUDA.jl-generated broadcast / mapreduce kernel, no KernelAbstractions kernel and no
# cuSPARSE / cuDSS call is left on the path of `mpc!` for a model whose vectors are CuVectors:
#
#   reference hook (file:line)                                        -> C ABI entry
#   MadIPM.coo_to_csr                 cuda_wrapper.jl:96-106           -> mipm_coo_to_csr
#   MadIPM.build_normal_system        cuda_wrapper.jl:214-234          -> mipm_normal_symbolic
#   linear_solver(aug_com; opt)       normalkkt.jl:113-115             -> B200Solver (mipm_ls_analyze)
#   MadNLP.factorize!/solve!          linear_solver.jl:10, normalkkt.jl:210 -> mipm_ls_factorize_async / mipm_ls_solve
#   MadIPM.is_factorized              src/utils.jl:54-62               -> mipm_ls_status
#   MadNLP.compress_jacobian!         cuda_wrapper.jl:32-41            -> mipm_gather + mipm_spmv_cache_values + mipm_normal_set_jacobian
#   MadNLP.build_kkt!(::NormalKKT)    normalkkt.jl:180-194             -> mipm_normal_assemble
#   MadNLP.transfer! (K2)             cuda_wrapper.jl:12-24            -> mipm_k2_transfer
#   MadNLP.jtprod!                    normalkkt.jl:176-178             -> mipm_spmv
#   MadNLP.solve!(::NormalKKT, w)     normalkkt.jl:196-219             -> mipm_normal_solve_stage x3 + mipm_spmv x2 + mipm_ls_solve
#   MadNLP.mul!(w, ::NormalKKT, v)    normalkkt.jl:221-233             -> mipm_spmv_pair + mipm_kktmul
#   MadIPM.solve_system!              linear_solver.jl:19-44           -> the above + mipm_residual_norms
#   MadIPM.init_starting_point!       solver.jl:6-125                  -> mipm_init_point_stage (+ fills / axpbys)
#   MadIPM.set_aug_diagonal_reg!      kernels.jl:124-149               -> mipm_set_aug_diagonal_reg[_scaled]
#   src/kernels.jl vector functions                                    -> mipm_set_* / mipm_get_* (fused)
#   MadIPM.update_step!(::MehrotraAdaptiveStep) kernels.jl:309-358     -> mipm_mehrotra_adaptive_step (no scalar indexing)
#   MadIPM.mpc!                       solver.jl:332-360                -> mipm_mpc_iter_begin / _peek / _refactor / _iter_rest (one sync per iteration)
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

# ---------------------------------------------------------------------------------------
# Handle: ONE per KKT system. It is created by whichever constructor runs first (the symbolic normal-matrix builder
# for NormalKKTSystem, the linear-solver constructor for K2 / K2.5), owned by the B200Solver and shared by every
# override below, because the library keeps the product-term map, the SpMV index, the factorization and the bound
# vectors of one problem in one handle (madipm_jl_b200/solver.py does the same).
# ---------------------------------------------------------------------------------------
mutable struct Handle
    ptr::Ptr{Cvoid}
    spmv_ready::Bool
    bound::Bool
    model_set::Bool
    function Handle(dev::Integer = CUDA.deviceid(CUDA.device()), stream = CUDA.stream())
        out = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:mipm_create, libmadipm), Cint, (Ref{Ptr{Cvoid}}, Cint, Ptr{Cvoid}),
                   out, dev, reinterpret(Ptr{Cvoid}, stream.handle))
        rc == MIPM_OK || error("mipm_create failed with code $rc (no CPU fallback exists)")
        h = new(out[], false, false, false)
        finalizer(x -> ccall((:mipm_destroy, libmadipm), Cint, (Ptr{Cvoid},), x.ptr), h)
        return h
    end
end

struct B200Error <: Exception
    code::Cint
    msg::String
end
Base.showerror(io::IO, e::B200Error) = print(io, "libmadipm_b200 error ", e.code, ": ", e.msg)

# Non-zero return codes carry the library's message (mipm_last_error). Pivot breakdown is NOT an error code: it is
# reported through mipm_ls_status and becomes is_factorized(ls) == false (reference retry path, linear_solver.jl:6-17).
function check(h::Handle, rc::Cint)
    rc == MIPM_OK && return
    msg = unsafe_string(ccall((:mipm_last_error, libmadipm), Cstring, (Ptr{Cvoid},), h.ptr))
    throw(B200Error(rc, msg))
end

devptr(x::CuArray) = reinterpret(Ptr{Cvoid}, pointer(x))
devptr(x::SubArray{<:Any,1,<:CuArray}) = reinterpret(Ptr{Cvoid}, pointer(x))   # contiguous views only (primal / dual blocks)

# The handle created by build_normal_system is adopted by the linear-solver constructor that runs right after it in
# create_kkt_system (normalkkt.jl:95-115); (m, nnzC) identify the matrix it was built for.
const PENDING = Ref{Any}(nothing)

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

const SOLVER_OF_CSC = IdDict{Any,Any}()          # aug_com => B200Solver (K2: MadNLP.transfer! only sees the matrix)

function B200Solver(csc::CUSPARSE.CuSparseMatrixCSC{T,Int32};
                    opt = B200Options(), logger = MadNLP.MadNLPLogger()) where {T}
    n = size(csc, 1)
    pend = PENDING[]
    h = if pend !== nothing && pend.m == n && pend.nnz == length(csc.nzVal)
        PENDING[] = nothing
        pend.handle                      # carries the product-term map of this very matrix
    else
        Handle()
    end
    colptr = Vector(csc.colPtr)          # host copies for the one-time analysis, 1-based Int32
    rowval = Vector(csc.rowVal)
    kind = opt.b200_algorithm == MadNLP.CHOLESKY ? MIPM_CHOLESKY : MIPM_LDL
    check(h, ccall((:mipm_ls_analyze, libmadipm), Cint,
                   (Ptr{Cvoid}, Int64, Ptr{Int32}, Ptr{Int32}, Cint, Cint, Cint, Ptr{Int32}),
                   h.ptr, n, colptr, rowval, 1, kind, opt.b200_ordering, C_NULL))
    M = B200Solver{T}(h, csc, opt, logger)
    SOLVER_OF_CSC[csc] = M
    return M
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
const GPUNormalKKT{T} = MadIPM.NormalKKTSystem{T,VT,MT} where {VT,MT<:CUSPARSE.CuSparseMatrixCSC{T,Int32}}

kkt_handle(kkt) = kkt.linear_solver.handle

# CSR of A = colptr / rowval of AT (normalkkt.jl:92), registered once for every SpMV of the solve.
function ensure_spmv(kkt::GPUNormalKKT)
    h = kkt_handle(kkt)
    h.spmv_ready && return h
    Ap, Aj = Vector(kkt.AT.colPtr), Vector(kkt.AT.rowVal)
    check(h, ccall((:mipm_spmv_setup, libmadipm), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Int32}, Ptr{Int32}, Cint),
                   h.ptr, kkt.m, kkt.n, Ap, Aj, 1))
    h.spmv_ready = true
    return h
end

# COO -> CSR with the value map (host arrays; normalkkt.jl:84-88 pushes V = 1:nnz through it).
function MadIPM.coo_to_csr(n_rows, n_cols, Ai::Vector{Int32}, Aj::Vector{Int32}, Ax::Vector{Tv}) where {Tv}
    nnz = length(Ai)
    Bp, Bj, Bmap = Vector{Int32}(undef, n_rows + 1), Vector{Int32}(undef, nnz), Vector{Int64}(undef, nnz)
    rc = ccall((:mipm_coo_to_csr, libmadipm), Cint,
               (Int64, Int64, Int64, Ptr{Int32}, Ptr{Int32}, Cint, Ptr{Int32}, Ptr{Int32}, Ptr{Int64}),
               n_rows, n_cols, nnz, Ai, Aj, 1, Bp, Bj, Bmap)
    rc == MIPM_OK || throw(B200Error(rc, "mipm_coo_to_csr"))
    return Bp, Bj, Ax[Bmap]
end

# Symbolic tril(A A'): host arrays in, host arrays out (normalkkt.jl:104 calls it with Vectors). The handle that now
# holds the product-term map is parked in PENDING and adopted by the B200Solver constructed from the returned pattern.
function MadIPM.build_normal_system(n_rows, n_cols, Jtp::Vector{Int32}, Jtj::Vector{Int32})
    h = Handle()
    Cp, Cj, nnzC = Ref{Ptr{Int32}}(C_NULL), Ref{Ptr{Int32}}(C_NULL), Ref{Int64}(0)
    check(h, ccall((:mipm_normal_symbolic, libmadipm), Cint,
                   (Ptr{Cvoid}, Int64, Int64, Ptr{Int32}, Ptr{Int32}, Cint, Ref{Ptr{Int32}}, Ref{Ptr{Int32}}, Ref{Int64}),
                   h.ptr, n_rows, n_cols, Jtp, Jtj, 1, Cp, Cj, nnzC))
    cp = copy(unsafe_wrap(Array, Cp[], n_rows + 1)); ccall((:mipm_free, libmadipm), Cvoid, (Ptr{Cvoid},), Cp[])
    cj = copy(unsafe_wrap(Array, Cj[], nnzC[]));     ccall((:mipm_free, libmadipm), Cvoid, (Ptr{Cvoid},), Cj[])
    PENDING[] = (handle = h, m = n_rows, nnz = Int(nnzC[]))
    return (cp, cj)
end

function MadNLP.compress_jacobian!(kkt::GPUNormalKKT{T}) where {T}
    h = ensure_spmv(kkt)
    n_slack = length(kkt.ind_ineq)
    n_slack > 0 && check(h, ccall((:mipm_fill, libmadipm), Cint, (Ptr{Cvoid}, Int64, Cdouble, Ptr{Cvoid}),
                                  h.ptr, n_slack, -1.0, devptr(view(kkt.A.V, length(kkt.A.V)-n_slack+1:length(kkt.A.V)))))
    # AT.nzVal .= A.V[A_csr_map] (cuda_wrapper.jl:39): one device gather through the Int64 map, once per solve
    check(h, ccall((:mipm_gather, libmadipm), Cint, (Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Cvoid}),
                   h.ptr, length(kkt.A_csr_map), devptr(kkt.A.V), devptr(kkt.A_csr_map), 1, devptr(kkt.AT.nzVal)))
    check(h, ccall((:mipm_spmv_cache_values, libmadipm), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), h.ptr, devptr(kkt.AT.nzVal)))
    check(h, ccall((:mipm_normal_set_jacobian, libmadipm), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), h.ptr, devptr(kkt.AT.nzVal)))
    return
end

# build_kkt!(::NormalKKTSystem): D = 1 ./ pr_diag and the assembly in one call (normalkkt.jl:180-194).
function MadNLP.build_kkt!(kkt::GPUNormalKKT)
    h = kkt_handle(kkt)
    check(h, ccall((:mipm_normal_assemble, libmadipm), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint),
                   h.ptr, devptr(kkt.pr_diag), devptr(kkt.aug_com.nzVal), 0))
    return
end

spmv!(h::Handle, trans, alpha, Ax, x, beta, y) =
    check(h, ccall((:mipm_spmv, libmadipm), Cint, (Ptr{Cvoid}, Cint, Cdouble, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}),
                   h.ptr, trans, alpha, devptr(Ax), devptr(x), beta, devptr(y)))

# jtprod!(y, kkt, x) = A' x (normalkkt.jl:176-178)
function MadNLP.jtprod!(y::CuVector, kkt::GPUNormalKKT, x::CuVector)
    spmv!(ensure_spmv(kkt), 1, 1.0, kkt.AT.nzVal, x, 0.0, y)
    return y
end

normal_stage!(h::Handle, stage, w, kkt) =
    check(h, ccall((:mipm_normal_solve_stage, libmadipm), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                   h.ptr, stage, devptr(MadNLP.full(w)), devptr(kkt.buffer_n), devptr(kkt.buffer_m)))

# solve!(kkt, w), normalkkt.jl:196-219: reduce_rhs! + r1 = wx ./ Sigma, r2 = wy | r2 = A r1 - r2 | dy = C \ r2 |
# wy = dy, r1 = wx | r1 -= A' wy | wx = r1 ./ Sigma + finish_aug_solve!  (needs mipm_mpc_bind: solver_handle)
function MadNLP.solve!(kkt::GPUNormalKKT, w::MadNLP.AbstractKKTVector)
    h = ensure_spmv(kkt)
    normal_stage!(h, 0, w, kkt)
    spmv!(h, 0, 1.0, kkt.AT.nzVal, kkt.buffer_n, -1.0, kkt.buffer_m)
    MadNLP.solve!(kkt.linear_solver, kkt.buffer_m)
    normal_stage!(h, 1, w, kkt)
    spmv!(h, 1, -1.0, kkt.AT.nzVal, kkt.buffer_m, 1.0, kkt.buffer_n)
    normal_stage!(h, 2, w, kkt)
    return w
end

# mul!(w, kkt, v, alpha, beta), normalkkt.jl:221-233
function LinearAlgebra.mul!(w::MadNLP.AbstractKKTVector{T}, kkt::GPUNormalKKT, v::MadNLP.AbstractKKTVector,
                            alpha = one(T), beta = zero(T)) where {T}
    h = ensure_spmv(kkt)
    # A v_x -> w_y and A' v_y -> w_x are independent: one launch
    check(h, ccall((:mipm_spmv_pair, libmadipm), Cint,
                   (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}),
                   h.ptr, devptr(kkt.AT.nzVal), alpha, devptr(MadNLP.primal(v)), beta, devptr(MadNLP.dual(w)),
                   alpha, devptr(MadNLP.dual(v)), beta, devptr(MadNLP.primal(w))))
    check(h, ccall((:mipm_kktmul, libmadipm), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cdouble),
                   h.ptr, devptr(MadNLP.full(w)), devptr(MadNLP.full(v)), alpha, beta))
    return w
end

# K2: MadNLP.transfer!(aug_com, aug_raw, aug_csc_map) as a deterministic gather (cuda_wrapper.jl:12-24 races on
# duplicates). The slot lists are rebuilt once, in the handle of the solver that owns aug_com, through mipm_k2_symbolic
# (same pattern as MadNLP's coo_to_csc; the returned host arrays are not needed and are freed).
const K2_READY = IdDict{Any,Bool}()
function MadNLP.transfer!(dest::CUSPARSE.CuSparseMatrixCSC{Tv}, src::MadNLP.SparseMatrixCOO{Tv}, map::CuVector{Int}) where {Tv}
    M = get(SOLVER_OF_CSC, dest, nothing)
    M === nothing && error("MadNLP.transfer!: the destination is not the matrix of a B200Solver")
    h = M.handle
    if !get(K2_READY, dest, false)
        colptr, rowval, m, nnz = Ref{Ptr{Int32}}(C_NULL), Ref{Ptr{Int32}}(C_NULL), Ref{Ptr{Int64}}(C_NULL), Ref{Int64}(0)
        I, J = Vector{Int32}(src.I), Vector{Int32}(src.J)
        check(h, ccall((:mipm_k2_symbolic, libmadipm), Cint,
                       (Ptr{Cvoid}, Int64, Int64, Ptr{Int32}, Ptr{Int32}, Cint, Ref{Ptr{Int32}}, Ref{Ptr{Int32}}, Ref{Ptr{Int64}}, Ref{Int64}),
                       h.ptr, size(dest, 1), length(I), I, J, 1, colptr, rowval, m, nnz))
        for p in (colptr[], rowval[], m[]); ccall((:mipm_free, libmadipm), Cvoid, (Ptr{Cvoid},), p); end
        K2_READY[dest] = true
    end
    check(h, ccall((:mipm_k2_transfer, libmadipm), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                   h.ptr, devptr(src.V), devptr(nonzeros(dest))))
    return
end

# ---------------------------------------------------------------------------------------
# Fused vector kernels: specialise src/kernels.jl, src/linear_solver.jl and src/solver.jl on GPU solvers
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

# The handle of the solver's KKT system with the solver's vectors bound to it (once).
function solver_handle(s::GPUSolver)
    h = kkt_handle(s.kkt)
    h.bound && return h
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
    h.bound = true
    return h
end

function MadIPM.set_aug_diagonal_reg!(kkt::MadNLP.AbstractKKTSystem{T}, s::GPUSolver{T}) where {T}
    h = solver_handle(s)
    check(h, ccall((:mipm_set_aug_diagonal_reg, libmadipm), Cint, (Ptr{Cvoid}, Cdouble, Cdouble), h.ptr, s.del_w, s.del_c))
end
# K2.5 (kernels.jl:139-149): positive l_diag / u_diag, pr_diag and the scaling factor in one launch
function MadIPM.set_aug_diagonal_reg!(kkt::MadNLP.ScaledSparseKKTSystem{T}, s::GPUSolver{T}) where {T}
    h = solver_handle(s)
    check(h, ccall((:mipm_set_aug_diagonal_reg_scaled, libmadipm), Cint, (Ptr{Cvoid}, Cdouble, Cdouble, Ptr{Cvoid}),
                   h.ptr, s.del_w, s.del_c, devptr(kkt.scaling_factor)))
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
# MehrotraAdaptiveStep without scalar indexing into device arrays (kernels.jl:309-358 reads single elements from the host)
function MadIPM.update_step!(rule::MadIPM.MehrotraAdaptiveStep, s::GPUSolver)
    h = solver_handle(s); a = zeros(Cdouble, 2)
    check(h, ccall((:mipm_mehrotra_adaptive_step, libmadipm), Cint, (Ptr{Cvoid}, Cdouble, Ptr{Cdouble}), h.ptr, rule.gamma_f, a))
    s.alpha_p, s.alpha_d = a[1], a[2]
    return
end
function MadIPM.apply_step!(s::GPUSolver)
    h = solver_handle(s)
    check(h, ccall((:mipm_apply_step, libmadipm), Cint, (Ptr{Cvoid}, Cdouble, Cdouble, Cdouble), h.ptr, s.alpha_p, s.alpha_d, s.mu))
    s.cnt.k += 1
    return
end

# update_termination_criteria! reductions (solver.jl:194-205, kernels.jl:408-430) in one launch, one sync
function termination_measures(s::GPUSolver)
    h = solver_handle(s); out = zeros(Cdouble, 5)
    check(h, ccall((:mipm_termination_measures, libmadipm), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), h.ptr, out))
    return out          # (dual objective, ||c||inf, ||f - zl + zu + jacl||inf, max complementarity, ||dx||inf)
end

# The status tests of update_termination_criteria! (solver.jl:206-222) on host scalars.
function termination_status!(s::GPUSolver, dobj)
    s.best_complementarity = min(s.best_complementarity, s.inf_compl)
    if max(s.inf_pr, s.inf_du, s.inf_compl) <= s.opt.tol
        s.status = MadNLP.SOLVE_SUCCEEDED
    elseif (s.inf_compl > s.opt.divergence_tol * s.best_complementarity) && (dobj > max(10.0 * abs(s.obj_val), 1.0))
        s.status = MadNLP.INFEASIBLE_PROBLEM_DETECTED
    elseif s.obj_val < -s.opt.divergence_tol * max(10.0, abs(dobj), 1.0)
        s.status = MadNLP.DIVERGING_ITERATES
    elseif s.cnt.k >= s.opt.max_iter
        s.status = MadNLP.MAXIMUM_ITERATIONS_EXCEEDED
    elseif time() - s.cnt.start_time >= s.opt.max_wall_time
        s.status = MadNLP.MAXIMUM_WALLTIME_EXCEEDED
    end
    return
end

# update_termination_criteria!, solver.jl:194-222: one fused launch instead of ~7 reductions
function MadIPM.update_termination_criteria!(s::GPUSolver)
    t = termination_measures(s)
    s.inf_pr = t[2] / max(1.0, s.norm_b)
    s.inf_du = t[3] / max(1.0, s.norm_c)
    s.inf_compl = t[4] / max(1.0, s.norm_c)
    termination_status!(s, t[1])
    return
end
MadIPM.dual_objective(s::GPUSolver) = termination_measures(s)[1]

# solve_system!, linear_solver.jl:19-44: d = K \ p, w = p - K d, residual ratio from one fused norm launch
function MadIPM.solve_system!(d::MadNLP.UnreducedKKTVector{T}, s::GPUSolver{T}, p::MadNLP.UnreducedKKTVector{T}) where {T}
    h = solver_handle(s)
    N = length(MadNLP.full(d))
    copyfull!(dst, src) = check(h, ccall((:mipm_copy, libmadipm), Cint, (Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}),
                                         h.ptr, N, devptr(MadNLP.full(src)), devptr(MadNLP.full(dst))))
    copyfull!(d, p)
    MadNLP.solve!(s.kkt, d)
    w = s._w1
    copyfull!(w, p)
    mul!(w, s.kkt, d, -one(T), one(T))
    nrm = zeros(Cdouble, 2)
    check(h, ccall((:mipm_residual_norms, libmadipm), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cdouble}),
                   h.ptr, devptr(MadNLP.full(w)), devptr(MadNLP.full(p)), nrm))
    residual_ratio = nrm[1] / max(one(T), nrm[2])
    if isnan(residual_ratio) || (s.opt.check_residual && (residual_ratio > s.opt.tol_linear_solve))
        throw(MadNLP.SolveException)
    end
    return d
end

# init_starting_point!, solver.jl:6-125: the two least-squares solves stay the reference's calls (they dispatch to the
# overrides above); the ~20 broadcast / mapreduce statements after them are three fused launches.
function MadIPM.init_starting_point!(s::GPUSolver{T}) where {T}
    h = solver_handle(s)
    fill!(v, a) = check(h, ccall((:mipm_fill, libmadipm), Cint, (Ptr{Cvoid}, Int64, Cdouble, Ptr{Cvoid}), h.ptr, length(v), a, devptr(v)))
    axpby!(n, a, x, b, y) = check(h, ccall((:mipm_axpby, libmadipm), Cint, (Ptr{Cvoid}, Int64, Cdouble, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}),
                                            h.ptr, n, a, devptr(x), b, devptr(y)))
    stage(k, a, b, kappa, cnt) = begin
        out = zeros(Cdouble, 5)
        check(h, ccall((:mipm_init_point_stage, libmadipm), Cint, (Ptr{Cvoid}, Cint, Cdouble, Cdouble, Cdouble, Ptr{Cdouble}),
                       h.ptr, k, a, b, kappa, out))
        out[1:cnt]
    end
    fill!(s.kkt.reg, s.del_w); fill!(s.kkt.pr_diag, s.del_w); fill!(s.kkt.du_diag, s.del_c)
    MadNLP.factorize_wrapper!(s)
    # Step 1: p = [0; -c; 0; 0]  (set_initial_primal_rhs!, kernels.jl:1-9)
    fill!(MadNLP.full(s.p), 0.0)
    axpby!(s.m, -1.0, s.c, 0.0, MadNLP.dual(s.p))
    MadIPM.solve_system!(s.d, s, s.p)
    axpby!(s.n, 1.0, MadNLP.primal(s.d), 1.0, MadNLP.full(s.x))
    # Step 2: p = [-f; 0; 0; 0]  (set_initial_dual_rhs!, kernels.jl:11-19)
    fill!(MadNLP.full(s.p), 0.0)
    axpby!(s.n, -1.0, MadNLP.full(s.f), 0.0, MadNLP.primal(s.p))
    MadIPM.solve_system!(s.d, s, s.p)
    check(h, ccall((:mipm_copy, libmadipm), Cint, (Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}), h.ptr, s.m, devptr(MadNLP.dual(s.d)), devptr(s.y)))
    # Step 3: res = A'y + f held in jacl, then the three fused stages
    MadNLP.jtprod!(s.jacl, s.kkt, s.y)
    axpby!(s.n, 1.0, MadNLP.full(s.f), 1.0, s.jacl)
    mins = stage(0, 0.0, 0.0, 0.0, 4)
    delta_x = max(0.0, -1.5 * mins[1], -1.5 * mins[2])
    delta_s = max(0.0, -1.5 * mins[3], -1.5 * mins[4])
    sm = stage(1, delta_x, delta_s, 0.0, 5)
    delta_x2 = sm[1] / (2 * (sm[2] + sm[3]))
    delta_s2 = sm[1] / (2 * (sm[4] + sm[5]))
    chk = stage(2, delta_x2, delta_s2, s.opt.bound_fac, 4)
    @assert s.nlb == 0 || (chk[1] > 0.0 && chk[3] > 0.0)
    @assert s.nub == 0 || (chk[2] > 0.0 && chk[4] > 0.0)
    return
end

# ---------------------------------------------------------------------------------------
# Fused mpc! (solver.jl:332-360): step lengths, centering parameter and barrier value stay on the device; ONE host
# synchronisation per iteration (termination measures + factorization status). Used for the default options
# (no Gondzio corrections, no residual check); otherwise the reference loop runs with the overrides above.
# ---------------------------------------------------------------------------------------
struct MpcModel
    kkt_kind::Cint; exact_order::Cint
    nx::Int64; c0::Cdouble
    ATx::Ptr{Cvoid}; cvec::Ptr{Cvoid}; Hx::Ptr{Cvoid}
    aug_nz::Ptr{Cvoid}; aug_raw_V::Ptr{Cvoid}
    buffer_n::Ptr{Cvoid}; buffer_m::Ptr{Cvoid}
end

step_rule_code(r::MadIPM.AdaptiveStep) = (Cint(0), r.tau_min)
step_rule_code(r::MadIPM.ConservativeStep) = (Cint(1), r.tau)
step_rule_code(r::MadIPM.MehrotraAdaptiveStep) = (Cint(2), r.gamma_f)

use_fused(s::GPUSolver) = s.opt.max_ncorr <= 0 && !s.opt.check_residual && s.kkt isa GPUNormalKKT

function MadIPM.mpc!(s::GPUSolver{T}) where {T}
    use_fused(s) || return invoke(MadIPM.mpc!, Tuple{MadNLP.AbstractMadNLPSolver}, s)
    h = solver_handle(s)
    kkt = s.kkt
    if !h.model_set
        qp = s.nlp
        md = MpcModel(0, 0, qp.meta.nvar, s.cb.obj_scale[] * qp.data.c0, devptr(kkt.AT.nzVal), devptr(qp.data.c), C_NULL,
                      devptr(kkt.aug_com.nzVal), C_NULL, devptr(kkt.buffer_n), devptr(kkt.buffer_m))
        check(h, ccall((:mipm_mpc_set_model, libmadipm), Cint, (Ptr{Cvoid}, Ref{MpcModel}), h.ptr, Ref(md)))
        h.model_set = true
    end
    out, st = zeros(Cdouble, 16), Ref{Cint}(0)
    rule, tau = step_rule_code(s.opt.step_rule)
    started = false
    while true
        MadNLP.print_iter(s)
        MadIPM.update_regularization!(s, s.opt.regularization)
        # close to convergence the measures are read before the next system is factorized (mipm_mpc_peek), so the call
        # that detects convergence does not pay for a factorization nobody uses
        peek = started && max(s.inf_pr, s.inf_du, s.inf_compl) <= 10 * s.opt.tol
        if peek
            check(h, ccall((:mipm_mpc_peek, libmadipm), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), h.ptr, out))
        else
            check(h, ccall((:mipm_mpc_iter_begin, libmadipm), Cint, (Ptr{Cvoid}, Cdouble, Cdouble, Ptr{Cdouble}, Ref{Cint}),
                           h.ptr, s.del_w, s.del_c, out, st))
        end
        if started                                    # scalars of the step taken by the previous mipm_mpc_iter_rest
            s.obj_val = s.cb.obj_scale[] * s.nlp.data.c0 + out[6] + 0.5 * out[7]
            s.alpha_p, s.alpha_d, s.mu, s.mu_curr = out[8], out[9], out[10], out[11]
            (isnan(out[12]) || isnan(out[14])) && throw(MadNLP.SolveException)
        end
        # update_termination_criteria! (solver.jl:194-222) from the fused reductions
        s.inf_pr = out[2] / max(1.0, s.norm_b)
        s.inf_du = out[3] / max(1.0, s.norm_c)
        s.inf_compl = out[4] / max(1.0, s.norm_c)
        termination_status!(s, out[1])
        MadIPM.is_done(s) && return
        if peek
            check(h, ccall((:mipm_mpc_iter_begin, libmadipm), Cint, (Ptr{Cvoid}, Cdouble, Cdouble, Ptr{Cdouble}, Ref{Cint}),
                           h.ptr, s.del_w, s.del_c, out, st))
        end
        ok = st[] == MIPM_OK
        ntrial = 1
        while !ok && ntrial < 3                       # factorize_regularized_system! (linear_solver.jl:6-17)
            s.del_w *= 100.0; s.del_c *= 100.0
            check(h, ccall((:mipm_mpc_refactor, libmadipm), Cint, (Ptr{Cvoid}, Cdouble, Cdouble, Ref{Cint}), h.ptr, s.del_w, s.del_c, st))
            ok = st[] == MIPM_OK
            ntrial += 1
        end
        check(h, ccall((:mipm_mpc_iter_rest, libmadipm), Cint, (Ptr{Cvoid}, Cdouble, Cint, Cdouble, Cint),
                       h.ptr, s.opt.mu_min, rule, tau, 0))
        s.cnt.k += 1
        started = true
    end
end

end # module
