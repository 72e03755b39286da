# dump_reference_phi.jl — writes the UNMODIFIED reference's `SLS_𝓗₂` output for BASELINE config #1 (the README chain,
# README.md:43-57: Nx = 59, Nu = 20, d = 9, T = 29, α = 1.5) to a plain-text file that this repository's tests consume:
#
#   julia --project=/path/to/SystemLevelControl.jl baseline/dump_reference_phi.jl tests/golden/reference_phi_config1.txt
#   python -m pytest tests/test_reference_pin.py            # (CPU: oracle vs the file;  -m gpu: CUDA library vs the file)
#
# Not runnable in this repository's image (no `julia`; JuMP 1.10.0 / Ipopt 3.14.10 / MUMPS 5.5.1 are not vendored with the
# reference).  It is the one-command route from "Φ values parity-unpinned" (DESIGN.md §5) to a reference-pinned oracle:
# commit the file it writes and tests/test_reference_pin.py stops skipping.  Only Julia's stdlib is used for the output.
#
# File format (text, 1-based indices as Julia stores them):
#   SLSPHI 1 <Nx> <Nu> <T>
#   then for fam in (x, u), t in 1:T:   "<fam> <t> <nrow> <ncol> <nnz>"  followed by three lines: colptr, rowval, nzval (%.17g)
using SystemLevelControl, LinearAlgebra, SparseArrays, Printf

out = length(ARGS) >= 1 ? ARGS[1] : "reference_phi_config1.txt"
Nx = 59
A  = I + spdiagm(1 => 0.2ones(Nx-1)) - spdiagm(-1 => 0.2ones(Nx-1))          # README.md:45
B₁ = sparse(1.0I, Nx, Nx)
B₂ = sparse(1.0I, Nx, Nx)[:, vec((1:2) .+ 6(0:9)')]                           # README.md:47  (Nu = 20)
P  = Plant(A, B₁, B₂)
d, T, α = (9, 29, 1.5)                                                         # README.md:52
𝓢ₓ = [           (A .≠ 0)^Int(min(d,   floor(α*(t-1)))) .≠ 0 for t = 1:T]      # README.md:53
𝓢ᵤ = [(B₂' .≠ 0)*(A .≠ 0)^Int(min(d+1, floor(α*(t-1)))) .≠ 0 for t = 1:T]      # README.md:54
Φₓ, Φᵤ = SLS_𝓗₂(P, [𝓢ₓ, 𝓢ᵤ])                                                  # README.md:57

open(out, "w") do io
    @printf(io, "SLSPHI 1 %d %d %d\n", Nx, size(B₂, 2), T)
    for (fam, Φ) in (("x", Φₓ), ("u", Φᵤ)), t in 1:T
        M = SparseMatrixCSC{Float64,Int64}(Φ[t])
        @printf(io, "%s %d %d %d %d\n", fam, t, size(M, 1), size(M, 2), nnz(M))
        println(io, join(M.colptr, " "))
        println(io, join(M.rowval, " "))
        println(io, join((@sprintf("%.17g", v) for v in M.nzval), " "))
    end
end
println("wrote $(out)")
