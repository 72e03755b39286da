# run_reference.jl — times the UNMODIFIED reference (`SLS_𝓗₂`, JuMP + Ipopt) on BASELINE.json's workload.
#
#   julia -p <physical cores> --project=/path/to/SystemLevelControl.jl baseline/run_reference.jl [Nx] [ncols]
#
# Not runnable in this repository's image (no `julia`, and JuMP 1.10.0 / Ipopt 3.14.10 / MUMPS 5.5.1 are not
# vendored with the reference: Manifest.toml:173-183,202-206,249-253); `bench.py --impl reference` therefore times the
# CPU oracle port (oracle/sls_oracle.py) and says so in its JSON line.  This script is what a maintainer with Julia
# runs to obtain the true reference number for the same metric (columns/s on the scaled chain, d=9, T=29, α=1.5).
# The reference is Θ(Nx) per column (it recomputes the global product 𝓢[end]*(A.≠0) for every column,
# src/reduction.jl:14, and re-allocates Θ(Nx) sparse matrices per column, src/synthesis.jl:67), so it is timed on an
# evenly spaced subset of columns through the `𝓘` keyword (src/synthesis.jl:11,15) and reported per column.
using Distributed
@everywhere using SystemLevelControl, LinearAlgebra, SparseArrays

Nx    = length(ARGS) >= 1 ? parse(Int, ARGS[1]) : 100_000
ncols = length(ARGS) >= 2 ? parse(Int, ARGS[2]) : 1_000

# README.md:43-54 scaled to Nx states
A  = I + spdiagm(1 => 0.2ones(Nx-1)) - spdiagm(-1 => 0.2ones(Nx-1))
B₁ = sparse(1.0I, Nx, Nx)
act = [j for n in 0:cld(Nx, 6) for j in (6n+1, 6n+2) if j <= Nx]
B₂ = sparse(1.0I, Nx, Nx)[:, act]
P  = Plant(A, B₁, B₂)
d, T, α = (9, 29, 1.5)
𝓢ₓ = [           (A .≠ 0)^Int(min(d,   floor(α*(t-1)))) .≠ 0 for t = 1:T]
𝓢ᵤ = [(B₂' .≠ 0)*(A .≠ 0)^Int(min(d+1, floor(α*(t-1)))) .≠ 0 for t = 1:T]

𝓘 = [[j] for j in round.(Int, range(1, Nx, length=ncols))]
SLS_𝓗₂(P, [𝓢ₓ, 𝓢ᵤ]; 𝓘 = 𝓘[1:min(end, 2nworkers())])          # warm-up (compilation)
t = @elapsed Φₓ, Φᵤ = SLS_𝓗₂(P, [𝓢ₓ, 𝓢ᵤ]; 𝓘 = 𝓘)
println("""{"impl": "reference", "kind": "reference", "metric": "Phi synthesis throughput", "unit": "columns/s", """ *
        """"value": $(ncols / t), "cores": $(nworkers()), "sample": "$(ncols) columns evenly spaced over a $(Nx)-state chain, julia -p $(nworkers())", """ *
        """"seconds": $(t)}""")
