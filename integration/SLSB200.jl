# SLSB200.jl — the binding a SystemLevelControl.jl maintainer adds to route `SLS_𝓗₂` through libslsb200.
# Written against Julia 1.8 / SystemLevelControl v1.1.0-DEV; NOT executed in this repository's environment
# (no Julia in the image; ABI v6).  Every `ccall` targets a symbol declared in include/slsb200.h; the Python binding
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
                    cols::Union{Nothing,Vector{Int64}}; wx=nothing, wu=nothing, cost=nothing, grp_ptr=nothing, device::Integer=0)
    # cost = (C₁, D₁₂, D₁₁ or nothing, B₁ or nothing): general cost -> sls_set_cost (kernel 2g); grp_ptr::Vector{Int64}: 0-based
    # offsets of the column groups inside `cols` -> sls_set_groups
    Nx, Nu, T = size(A, 1), size(B₂, 2), length(𝓢ₓ)
    h = Handle(device=device)
    GC.@preserve A B₂ 𝓢ₓ 𝓢ᵤ cols wx wu cost grp_ptr begin
        if wx !== nothing || wu !== nothing        # before sls_set_plant: the device holds the similarity-scaled plant
            check(h.ptr, ccall((:sls_set_weights, LIB), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Ptr{Float64}),
                  h.ptr, Nx, Nu, wx === nothing ? C_NULL : wx, wu === nothing ? C_NULL : wu))
        end
        check(h.ptr, ccall((:sls_set_plant, LIB), Cint,
              (Ptr{Cvoid}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
              h.ptr, Nx, Nu, A.colptr, A.rowval, A.nzval, B₂.colptr, B₂.rowval, B₂.nzval))
        if cost !== nothing                        # after sls_set_plant: the cost matrices are sized by Nx, Nu
            C₁, D₁₂, D₁₁, B₁ = cost
            f3(M) = M === nothing ? (Ptr{Int64}(C_NULL), Ptr{Int64}(C_NULL), Ptr{Float64}(C_NULL)) : (pointer(M.colptr), pointer(M.rowval), pointer(M.nzval))
            check(h.ptr, ccall((:sls_set_cost, LIB), Cint,
                  (Ptr{Cvoid}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64},
                   Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
                  h.ptr, size(C₁, 1), B₁ === nothing ? (D₁₁ === nothing ? Nx : size(D₁₁, 2)) : size(B₁, 2),
                  f3(C₁)..., f3(D₁₂)..., f3(D₁₁)..., f3(B₁)...))
        end
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
        grp_ptr === nothing || check(h.ptr, ccall((:sls_set_groups, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Int64}),
                                                  h.ptr, length(grp_ptr) - 1, grp_ptr))
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


"""
    register!(h, arrays...) / unregister!(h, arrays...)

Page-lock ordinary Julia `Vector`s in place (`sls_host_register`): the copy pipelines of `sls_solve` / `sls_synthesize*`
only overlap with the kernels on page-locked memory.  Measured on config #3 (bench.py `e2e_pageable` / `e2e_registered`):
92.6 ms per call with pageable vectors, 34.7 ms once they are registered (registering 1.75 GB costs ~240 ms, once).
"""
register!(h::Handle, arrays...) = foreach(a -> isempty(a) || check(h.ptr, ccall((:sls_host_register, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), h.ptr, pointer(a), sizeof(a))), arrays)
unregister!(h::Handle, arrays...) = foreach(a -> isempty(a) || check(h.ptr, ccall((:sls_host_unregister, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}), h.ptr, pointer(a))), arrays)

"""
    synthesize_shard(h, A, B₂, 𝓢ₓ, 𝓢ᵤ, cols::UnitRange; halo=cols, values_only=true) -> (Φₓ, Φᵤ, status, resid)

One GPU's share of the `@distributed` loop (src/synthesis.jl:16,23-27) through `sls_synthesize_window`: only the support
columns `halo ⊇ cols ∪ {rows of nz(A[:,j]), j ∈ cols}` are uploaded (a `pointer(𝓢[t].colptr, first(halo))` view, no copy)
and the column SLAB Φ[t][:, cols] comes back (colptr of length(cols)+1 entries).  `values_only = true` is the documented
fast path: the stored pattern of Φ[t] is the mask pattern, a subset of 𝓢[t]'s; when the returned count equals
nnz(𝓢[t][:, cols]) the pattern IS 𝓢[t]'s, and the slab aliases a `view`-free copy of that rowval segment made on the host
only when needed (no 0.53 GB of Int64 row indices over PCIe at config #3: 27.6 ms instead of 35.6 ms).  Otherwise the
call is repeated with rowval outputs.
"""
function synthesize_shard(h::Handle, A::SparseMatrixCSC{Float64,Int64}, B₂::SparseMatrixCSC{Float64,Int64},
                          𝓢ₓ::Vector{SparseMatrixCSC{Bool,Int64}}, 𝓢ᵤ::Vector{SparseMatrixCSC{Bool,Int64}},
                          cols::UnitRange{Int}; halo::UnitRange{Int}=cols, values_only::Bool=true)
    Nx, Nu, T = size(A, 1), size(B₂, 2), length(𝓢ₓ)
    nc, w1, wn = length(cols), first(halo), length(halo)
    GC.@preserve A B₂ 𝓢ₓ 𝓢ᵤ begin
        check(h.ptr, ccall((:sls_set_plant, LIB), Cint,
              (Ptr{Cvoid}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
              h.ptr, Nx, Nu, A.colptr, A.rowval, A.nzval, B₂.colptr, B₂.rowval, B₂.nzval))
        xcp = [pointer(M.colptr, w1) for M in 𝓢ₓ]; ucp = [pointer(M.colptr, w1) for M in 𝓢ᵤ]     # window views
        xrv = [pointer(M.rowval) for M in 𝓢ₓ];     urv = [pointer(M.rowval) for M in 𝓢ᵤ]
        xnz = [Ptr{UInt8}(pointer(M.nzval)) for M in 𝓢ₓ]; unz = [Ptr{UInt8}(pointer(M.nzval)) for M in 𝓢ᵤ]
        seg(M) = M.colptr[first(cols)]:(M.colptr[last(cols) + 1] - 1)
        capx, capu = Int64[length(seg(M)) for M in 𝓢ₓ], Int64[length(seg(M)) for M in 𝓢ᵤ]
        cp() = [Vector{Int64}(undef, nc + 1) for _ in 1:T]
        xc, uc = cp(), cp()
        xv, uv = [Vector{Float64}(undef, capx[t]) for t in 1:T], [Vector{Float64}(undef, capu[t]) for t in 1:T]
        xr = values_only ? nothing : [Vector{Int64}(undef, capx[t]) for t in 1:T]
        ur = values_only ? nothing : [Vector{Int64}(undef, capu[t]) for t in 1:T]
        nnzx, nnzu = zeros(Int64, T), zeros(Int64, T)
        status, resid = zeros(Int32, nc), zeros(Float64, nc)
        p(v) = v === nothing ? C_NULL : [pointer(a) for a in v]
        GC.@preserve xc uc xv uv xr ur begin
            check(h.ptr, ccall((:sls_synthesize_window, LIB), Cint,
                  (Ptr{Cvoid}, Int64, Int64, Int64, Int64, Int64,
                   Ptr{Ptr{Int64}}, Ptr{Ptr{Int64}}, Ptr{Ptr{UInt8}}, Ptr{Ptr{Int64}}, Ptr{Ptr{Int64}}, Ptr{Ptr{UInt8}},
                   Int64, Int64, Ptr{Int64}, Ptr{Int64},
                   Ptr{Ptr{Int64}}, Ptr{Ptr{Int64}}, Ptr{Ptr{Float64}}, Ptr{Ptr{Int64}}, Ptr{Ptr{Int64}}, Ptr{Ptr{Float64}},
                   Ptr{Int64}, Ptr{Int64}, Ptr{Int32}, Ptr{Float64}),
                  h.ptr, T, w1, wn, 0, 0, xcp, xrv, xnz, ucp, urv, unz, first(cols), nc, capx, capu,
                  p(xc), p(xr), p(xv), p(uc), p(ur), p(uv), nnzx, nnzu, status, resid))
        end
        if values_only
            # pattern == support pattern?  (a stored `false` in 𝓢[t], or a support row outside s_x, shrinks the count)
            all(nnzx .== capx) && all(nnzu .== capu) ||
                return synthesize_shard(h, A, B₂, 𝓢ₓ, 𝓢ᵤ, cols; halo=halo, values_only=false)
            xr = [𝓢ₓ[t].rowval[seg(𝓢ₓ[t])] for t in 1:T]
            ur = [𝓢ᵤ[t].rowval[seg(𝓢ᵤ[t])] for t in 1:T]
        end
        Φₓ = [SparseMatrixCSC{Float64,Int64}(Nx, nc, xc[t], resize!(xr[t], nnzx[t]), resize!(xv[t], nnzx[t])) for t in 1:T]
        Φᵤ = [SparseMatrixCSC{Float64,Int64}(Nu, nc, uc[t], resize!(ur[t], nnzu[t]), resize!(uv[t], nnzu[t])) for t in 1:T]
        return Φₓ, Φᵤ, status, resid          # column slabs: hcat the ranks' slabs (or write them into one colptr-offset Φ[t])
    end
end

"""
    diagonal_cost(P) -> (wx, wu)  or  :general

The cost family the DMMA kernels solve (the Python mirror's `Plant.diagonal_cost`): `[C₁ D₁₂]` = blockdiag(Diagonal(wx),
Diagonal(wu)) > 0, `D₁₁ = 0`, diagonal `B₁` without zeros (a diagonal B₁ rescales each column's cost and does not change Φ).
Anything else returns `:general`: the caller then passes `cost = (C₁, D₁₂, D₁₁, B₁)` and the general kernel 2g solves it
(sls_set_cost; synthesis.jl:42-43,50,76-77).
"""
function diagonal_cost(P)
    Nx, Nu = P.Nx, P.Nu
    size(P.B₁) == (Nx, Nx) && isdiag(P.B₁) && all(!iszero, diag(P.B₁)) || return :general
    iszero(P.D₁₁) || return :general
    L = [P.C₁ P.D₁₂]
    size(L) == (Nx + Nu, Nx + Nu) && isdiag(L) && all(>(0), diag(L)) || return :general
    w = diag(L)
    all(==(1.0), w) ? (nothing, nothing) : (Vector{Float64}(w[1:Nx]), Vector{Float64}(w[Nx+1:end]))
end

"""
    groups_decouple(P, 𝓢ₓ, 𝓢ᵤ, 𝓘) -> Bool

Pattern check behind q > 1 groups (mirror of synthesis.py::_groups_decouple): every row that A or B₂ reaches from a member
column's own supports already lies inside that column's own s_x, so the group's union index set only adds empty rows.
"""
function groups_decouple(P, 𝓢ₓ, 𝓢ᵤ, 𝓘)
    pA, pB = P.A .≠ 0, P.B₂ .≠ 0
    own = (𝓢ₓ[end] * pA) .≠ 0
    reach = reduce((a, b) -> a .| b, [((pA * (S .≠ 0)) .≠ 0) for S in 𝓢ₓ]) .| reduce((a, b) -> a .| b, [((pB * (S .≠ 0)) .≠ 0) for S in 𝓢ᵤ])
    all(g -> length(g) < 2 || all(j -> issubset(findnz(reach[:, j])[1], findnz(own[:, j])[1]), g), 𝓘)
end

end # module

# --- drop-in replacement of SLS_𝓗₂ in src/synthesis.jl (signature and return shape unchanged) -----------------
# function SLS_𝓗₂(P::AbstractGeneralizedPlant, 𝓢::AbstractVector; 𝓘=nothing)
#     if P isa GeneralizedPlant{T,StateFeedback} where {T}
#         default = 𝓘 === nothing                                               # captured BEFORE the reassignment below
#         𝓘 = default ? [[i] for i in 1:P.Nx] : 𝓘;                             # :15
#         𝓢ₓ,𝓢ᵤ = 𝓢;  T = length(𝓢ₓ)                                           # :19-20
#         dc = SLSB200.diagonal_cost(P)                                         # (wx, wu) for the DMMA kernels, or :general
#         cols = Int64[c for g in 𝓘 for c in g]
#         # q > 1 groups: solved as ONE problem on the union index set (kernel 2g) unless the group provably equals its
#         # columns' problems (diagonal B₁ and a union that adds only empty constraint rows: groups_decouple).  List each
#         # group's columns ascending (I~ pairs the impulses with the columns by their order in s_x, reduction.jl:22-23).
#         coupled = any(g -> length(g) > 1, 𝓘) && (dc === :general || !SLSB200.groups_decouple(P, 𝓢ₓ, 𝓢ᵤ, 𝓘))
#         grp = coupled ? Int64[0; cumsum(length.(𝓘))] : nothing
#         if dc === :general || coupled                                         # weights travel inside C₁ / D₁₂ on this path
#             Φₓ,Φᵤ,status,resid = SLSB200.synthesize(P.A, P.B₂, 𝓢ₓ, 𝓢ᵤ, cols; grp_ptr=grp,
#                                      cost=(sparse(P.C₁), sparse(P.D₁₂), iszero(P.D₁₁) ? nothing : sparse(P.D₁₁), sparse(P.B₁)))
#         else
#             wx, wu = dc
#             Φₓ,Φᵤ,status,resid = default ? SLSB200.synthesize_pipelined(P.A, P.B₂, 𝓢ₓ, 𝓢ᵤ; wx=wx, wu=wu) :
#                                            SLSB200.synthesize(P.A, P.B₂, 𝓢ₓ, 𝓢ᵤ, cols; wx=wx, wu=wu)
#         end
#         any(!iszero, status .& 1) && @warn "SLS_𝓗₂: $(count(!iszero, status .& 1)) column(s) infeasible / not converged (status bit 0)"
#         # optional: dropzeros!.(Φₓ); dropzeros!.(Φᵤ) reproduces the zero-dropping of `Φ̃ += …` (:67)
#         Φ = [Φₓ Φᵤ]                                                           # T×2 Matrix{SparseMatrixCSC}
#         return eachcol(Φ)                                                      # :27
#     end
# end
#
# Multi-GPU with `julia -p N` (one worker per GPU, the @distributed shape of :23-27): worker r owns a contiguous column
# range and calls SLSB200.synthesize_shard(h, A, B₂, 𝓢ₓ, 𝓢ᵤ, cols; halo = (first(cols)-1):(last(cols)+1) ∩ 1:Nx) with its own
# handle (device = myid()-2); the slabs concatenate by column (`hcat`), which is all the (+) reduction of disjoint columns
# does.  To keep host work and PCIe bytes independent of the global Nx, a worker may pass the sub-plant A[pw,pw], B₂[pw,uw]
# (pw ⊇ every s_x of its columns) and the global supports with row_off_x = first(pw)-1, row_off_u = first(uw)-1
# (sls_synthesize_window's last two integer arguments before the arrays; bench.py does exactly this at N > 1).
#
# sparsity_dim_reduction(P, cⱼ, 𝓢) (src/reduction.jl:11-27) maps to sls_reduce:
#   ccall((:sls_reduce, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int64, Ptr{Int64}, Int64, Ref{Int64}, Ptr{Int64}, Int64, Ref{Int64}, Ptr{UInt8}),
#         h, cⱼ, length(cⱼ), sx, length(sx), nsx, su, length(su), nsu, iix)
#   then P̃ = view(P, (sₓ, [sₓ; P.Nx .+ sᵤ]), (sₓ, cⱼ, sᵤ)) and Ĩ = [I(P̃.Nx)[:,iiₓ] spzeros(P̃.Nx, P̃.Nw-sum(iiₓ))] as in :15,:23.
# The OutputFeedback branch (src/reduction.jl:16-18; 𝓢 = [𝓢ₓ, 𝓢ᵤ, 𝓢ᵧ]) maps to sls_reduce_of, which takes the LAST tap of each family:
#   cp = [S[end].colptr for S in 𝓢]; rv = [S[end].rowval for S in 𝓢]; nr = Int64[size(S[end],1) for S in 𝓢]
#   ccall((:sls_reduce_of, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Ptr{Int64}}, Ptr{Ptr{Int64}}, Ptr{Int64}, Int64, Ptr{Int64}, Int64,
#                                       Ptr{Ptr{Int64}}, Ptr{Int64}, Ptr{Int64}, Ptr{UInt8}),
#         h, 3, pointer.(cp), pointer.(rv), nr, size(𝓢[1][end],2), cⱼ, length(cⱼ), pointer.(s_out), Int64.(length.(s_out)), ns, iix)
#   then P̃ = view(P, (sₓ, [sₓ; P.Nx .+ sᵤ], sᵧ), (sₓ, cⱼ, sᵤ)) and Ĩ as above (:17,:22-23).  No OutputFeedback synthesis exists in the
#   reference (SLS_𝓗₂ returns nothing for such plants) and L⁺ ≡ 0 (src/synthesis.jl:21,52), so the hook has no ABI counterpart.
