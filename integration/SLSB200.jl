# SLSB200.jl — the binding a SystemLevelControl.jl maintainer adds to route `SLS_𝓗₂` through libslsb200.
# Written against Julia 1.8 / SystemLevelControl v1.1.0-DEV; NOT executed in this repository's environment
# (no Julia in the image).  Every `ccall` targets a symbol declared in include/slsb200.h; the Python binding
# (systemlevelcontrol.jl_b200/_capi.py) drives the same symbols and is what the tests exercise.
#
# Replaces: src/synthesis.jl:23-27 (the @distributed loop) and src/synthesis.jl:34-72 (_SLS_𝓗₂).
# Keeps verbatim: the SLS_𝓗₂ signature, the StateFeedback guard (:13), the default 𝓘 (:15),
# `𝓢ₓ,𝓢ᵤ = 𝓢; T = length(𝓢ₓ)` (:19-20) and the `eachcol(::T×2 Matrix{SparseMatrixCSC})` return (:27).
module SLSB200

using SparseArrays
const LIB = get(ENV, "SLSB200_LIB", "libslsb200")

struct Options
    device::Int32
    flags::Int32
    rank_tol::Float64
    feas_tol::Float64
end

mutable struct Handle
    ptr::Ptr{Cvoid}
end

last_error(h::Ptr{Cvoid}) = unsafe_string(ccall((:sls_last_error, LIB), Cstring, (Ptr{Cvoid},), h))
function check(h::Ptr{Cvoid}, rc::Cint)
    rc == 0 || error("libslsb200 error $(rc): $(last_error(h))")   # package convention: error(...) (GeneralizedPlant.jl:297-309)
    nothing
end

function Handle(; device::Integer=0, rank_tol::Real=0.0, feas_tol::Real=0.0)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    opts = Ref(Options(device, 0, rank_tol, feas_tol))
    rc = ccall((:sls_create, LIB), Cint, (Ref{Options}, Ref{Ptr{Cvoid}}), opts, out)
    check(C_NULL, rc)
    h = Handle(out[])
    finalizer(x -> ccall((:sls_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.ptr), h)
    h
end

"""
    synthesize(A, B₂, 𝓢ₓ, 𝓢ᵤ, cols; device=0) -> (Φₓ, Φᵤ, status, resid)

`A::SparseMatrixCSC{Float64,Int64}`, `B₂` likewise, `𝓢ₓ/𝓢ᵤ::Vector{SparseMatrixCSC{Bool,Int64}}`, `cols` the
1-based disturbance columns (`nothing` = 1:Nx).  All arrays are passed zero-copy (`colptr`, `rowval`, `nzval`).
"""
function synthesize(A::SparseMatrixCSC{Float64,Int64}, B₂::SparseMatrixCSC{Float64,Int64},
                    𝓢ₓ::Vector{SparseMatrixCSC{Bool,Int64}}, 𝓢ᵤ::Vector{SparseMatrixCSC{Bool,Int64}},
                    cols::Union{Nothing,Vector{Int64}}; device::Integer=0)
    Nx, Nu, T = size(A, 1), size(B₂, 2), length(𝓢ₓ)
    h = Handle(device=device)
    GC.@preserve A B₂ 𝓢ₓ 𝓢ᵤ cols begin
        check(h.ptr, ccall((:sls_set_plant, LIB), Cint,
              (Ptr{Cvoid}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
              h.ptr, Nx, Nu, A.colptr, A.rowval, A.nzval, B₂.colptr, B₂.rowval, B₂.nzval))
        ptrs(f, S) = [pointer(f(M)) for M in S]
        xcp, xrv, xnz = ptrs(M -> M.colptr, 𝓢ₓ), ptrs(M -> M.rowval, 𝓢ₓ), [Ptr{UInt8}(pointer(M.nzval)) for M in 𝓢ₓ]
        ucp, urv, unz = ptrs(M -> M.colptr, 𝓢ᵤ), ptrs(M -> M.rowval, 𝓢ᵤ), [Ptr{UInt8}(pointer(M.nzval)) for M in 𝓢ᵤ]
        check(h.ptr, ccall((:sls_set_supports, LIB), Cint,
              (Ptr{Cvoid}, Int64, Ptr{Ptr{Int64}}, Ptr{Ptr{Int64}}, Ptr{Ptr{UInt8}}, Ptr{Ptr{Int64}}, Ptr{Ptr{Int64}}, Ptr{Ptr{UInt8}}),
              h.ptr, T, xcp, xrv, xnz, ucp, urv, unz))
        ncols = cols === nothing ? Nx : length(cols)
        nnzx, nnzu = zeros(Int64, T), zeros(Int64, T)
        check(h.ptr, ccall((:sls_plan, LIB), Cint,
              (Ptr{Cvoid}, Ptr{Int64}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int32}, Ptr{Int32}),
              h.ptr, cols === nothing ? C_NULL : cols, ncols, nnzx, nnzu, C_NULL, C_NULL))
        # two-phase sizing: Julia owns every output buffer
        mk(nnz, nrow) = [SparseMatrixCSC{Float64,Int64}(nrow, Nx, ones(Int64, Nx + 1), Vector{Int64}(undef, nnz[t]), Vector{Float64}(undef, nnz[t])) for t in 1:T]
        Φₓ, Φᵤ = mk(nnzx, Nx), mk(nnzu, Nu)
        status, resid = zeros(Int32, ncols), zeros(Float64, ncols)
        GC.@preserve Φₓ Φᵤ begin
            check(h.ptr, ccall((:sls_solve, LIB), Cint,
                  (Ptr{Cvoid}, Ptr{Ptr{Int64}}, Ptr{Ptr{Int64}}, Ptr{Ptr{Float64}}, Ptr{Ptr{Int64}}, Ptr{Ptr{Int64}}, Ptr{Ptr{Float64}}, Ptr{Int32}, Ptr{Float64}),
                  h.ptr, ptrs(M -> M.colptr, Φₓ), ptrs(M -> M.rowval, Φₓ), ptrs(M -> M.nzval, Φₓ),
                  ptrs(M -> M.colptr, Φᵤ), ptrs(M -> M.rowval, Φᵤ), ptrs(M -> M.nzval, Φᵤ), status, resid))
        end
        return Φₓ, Φᵤ, status, resid
    end
end

"""
    synthesize_pipelined(A, B₂, 𝓢ₓ, 𝓢ᵤ; first_col=1, ncols=0, wx=nothing, wu=nothing, device=0)

One call for the default 𝓘 = 1:Nx (or a contiguous shard `first_col : first_col+ncols-1`): `sls_synthesize` overlaps the
host→device copy of the supports, kernels 1-3 and the device→host copy of Φ column chunk by column chunk.  Because the
stored pattern of Φ[t] is a subset of 𝓢[t]'s, outputs are allocated at nnz(𝓢[t]) and `resize!`d to the returned counts.
`wx`/`wu` are optional diagonal cost weights (`C₁ = [Diagonal(wx); 0]`, `D₁₂ = [0; Diagonal(wu)]`, sls_set_weights).
"""
function synthesize_pipelined(A::SparseMatrixCSC{Float64,Int64}, B₂::SparseMatrixCSC{Float64,Int64},
                              𝓢ₓ::Vector{SparseMatrixCSC{Bool,Int64}}, 𝓢ᵤ::Vector{SparseMatrixCSC{Bool,Int64}};
                              first_col::Integer=1, ncols::Integer=0, wx=nothing, wu=nothing, device::Integer=0)
    Nx, Nu, T = size(A, 1), size(B₂, 2), length(𝓢ₓ)
    nc = ncols <= 0 ? Nx : ncols
    h = Handle(device=device)
    GC.@preserve A B₂ 𝓢ₓ 𝓢ᵤ wx wu begin
        if wx !== nothing || wu !== nothing
            check(h.ptr, ccall((:sls_set_weights, LIB), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Ptr{Float64}),
                  h.ptr, Nx, Nu, wx === nothing ? C_NULL : wx, wu === nothing ? C_NULL : wu))
        end
        check(h.ptr, ccall((:sls_set_plant, LIB), Cint,
              (Ptr{Cvoid}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
              h.ptr, Nx, Nu, A.colptr, A.rowval, A.nzval, B₂.colptr, B₂.rowval, B₂.nzval))
        ptrs(f, S) = [pointer(f(M)) for M in S]
        xcp, xrv, xnz = ptrs(M -> M.colptr, 𝓢ₓ), ptrs(M -> M.rowval, 𝓢ₓ), [Ptr{UInt8}(pointer(M.nzval)) for M in 𝓢ₓ]
        ucp, urv, unz = ptrs(M -> M.colptr, 𝓢ᵤ), ptrs(M -> M.rowval, 𝓢ᵤ), [Ptr{UInt8}(pointer(M.nzval)) for M in 𝓢ᵤ]
        capx, capu = Int64[nnz(M) for M in 𝓢ₓ], Int64[nnz(M) for M in 𝓢ᵤ]
        mk(cap, nrow) = [SparseMatrixCSC{Float64,Int64}(nrow, Nx, ones(Int64, Nx + 1), Vector{Int64}(undef, cap[t]), Vector{Float64}(undef, cap[t])) for t in 1:T]
        Φₓ, Φᵤ = mk(capx, Nx), mk(capu, Nu)
        nnzx, nnzu = zeros(Int64, T), zeros(Int64, T)
        status, resid = zeros(Int32, nc), zeros(Float64, nc)
        GC.@preserve Φₓ Φᵤ begin
            check(h.ptr, ccall((:sls_synthesize, LIB), Cint,
                  (Ptr{Cvoid}, Int64, Ptr{Ptr{Int64}}, Ptr{Ptr{Int64}}, Ptr{Ptr{UInt8}}, Ptr{Ptr{Int64}}, Ptr{Ptr{Int64}}, Ptr{Ptr{UInt8}},
                   Int64, Int64, Ptr{Int64}, Ptr{Int64},
                   Ptr{Ptr{Int64}}, Ptr{Ptr{Int64}}, Ptr{Ptr{Float64}}, Ptr{Ptr{Int64}}, Ptr{Ptr{Int64}}, Ptr{Ptr{Float64}},
                   Ptr{Int64}, Ptr{Int64}, Ptr{Int32}, Ptr{Float64}),
                  h.ptr, T, xcp, xrv, xnz, ucp, urv, unz, first_col, ncols, capx, capu,
                  ptrs(M -> M.colptr, Φₓ), ptrs(M -> M.rowval, Φₓ), ptrs(M -> M.nzval, Φₓ),
                  ptrs(M -> M.colptr, Φᵤ), ptrs(M -> M.rowval, Φᵤ), ptrs(M -> M.nzval, Φᵤ), nnzx, nnzu, status, resid))
        end
        for t in 1:T
            resize!(Φₓ[t].rowval, nnzx[t]); resize!(Φₓ[t].nzval, nnzx[t])
            resize!(Φᵤ[t].rowval, nnzu[t]); resize!(Φᵤ[t].nzval, nnzu[t])
        end
        return Φₓ, Φᵤ, status, resid
    end
end

end # module

# --- drop-in replacement of SLS_𝓗₂ in src/synthesis.jl (signature and return shape unchanged) -----------------
# function SLS_𝓗₂(P::AbstractGeneralizedPlant, 𝓢::AbstractVector; 𝓘=nothing)
#     if P isa GeneralizedPlant{T,StateFeedback} where {T}
#         𝓘 = (𝓘 === nothing) ? [[i] for i in 1:P.Nx] : 𝓘;                      # :15
#         𝓢ₓ,𝓢ᵤ = 𝓢;  T = length(𝓢ₓ)                                           # :19-20
#         # q > 1 groups: with a diagonal B₁ the cost is separable and the group problem equals its columns' problems
#         # when the supports are closed under the plant's pattern (checked on the host; see synthesis.py::_groups_decouple)
#         cols = Int64[c for g in 𝓘 for c in g]
#         Φₓ,Φᵤ,status,resid = (𝓘 === nothing) ? SLSB200.synthesize_pipelined(P.A, P.B₂, 𝓢ₓ, 𝓢ᵤ) :
#                                                  SLSB200.synthesize(P.A, P.B₂, 𝓢ₓ, 𝓢ᵤ, cols)
#         # optional: dropzeros!.(Φₓ); dropzeros!.(Φᵤ) reproduces the zero-dropping of `Φ̃ += …` (:67)
#         Φ = [Φₓ Φᵤ]                                                           # T×2 Matrix{SparseMatrixCSC}
#         return eachcol(Φ)                                                      # :27
#     end
# end
#
# Multi-GPU with `julia -p N` (one worker per GPU, the @distributed shape of :23-27): each worker calls
# SLSB200.synthesize(...; device = myid()-2) on its slice of 𝓘; the (+) reduction concatenates disjoint columns.
#
# sparsity_dim_reduction(P, cⱼ, 𝓢) (src/reduction.jl:11-27) maps to sls_reduce:
#   ccall((:sls_reduce, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int64, Ptr{Int64}, Int64, Ref{Int64}, Ptr{Int64}, Int64, Ref{Int64}, Ptr{UInt8}),
#         h, cⱼ, length(cⱼ), sx, length(sx), nsx, su, length(su), nsu, iix)
#   then P̃ = view(P, (sₓ, [sₓ; P.Nx .+ sᵤ]), (sₓ, cⱼ, sᵤ)) and Ĩ = [I(P̃.Nx)[:,iiₓ] spzeros(P̃.Nx, P̃.Nw-sum(iiₓ))] as in :15,:23.
