# julia_reference_trace.jl -- NOT RUN in the build image (no Julia there); for anyone with Julia.
#
# Produces a GENUINE reference trace: runs the unmodified KissMCMC.emcee single-threaded under a
# seeded global RNG, then re-seeds and regenerates the draws in the order the reference consumes
# them (src/samplers.jl:250 partner `rand(ncos)`, :252/:230 `rand()` for z, :260 `rand()`), without
# copying any reference code.  Writes little-endian binaries that tests/test_julia_trace.py loads
# (the test is skipped while the files are absent) and replays through the CUDA path.
#
#   JULIA_NUM_THREADS=1 julia --project=/path/to/KissMCMC.jl oracle/julia_reference_trace.jl out_dir
using KissMCMC, Random

outdir = length(ARGS) >= 1 ? ARGS[1] : "tests/golden/julia"
mkpath(outdir)
@assert Threads.nthreads() == 1 "the draw order is only defined single-threaded"

rosen(x) = -(100 * (x[2] - x[1]^2)^2 + (1 - x[1])^2) / 20      # test/runtests.jl:68
nwalkers, niter, nburnin, nthin, a = 64, 64 * 60, 64 * 30, 2, 2.0
Random.seed!(20240601)
theta0s = [0.1 .* randn(2) for _ in 1:nwalkers]

Random.seed!(7)
thetas, accept_ratio, logdensities, _ = emcee(rosen, theta0s; niter=niter, nburnin=nburnin, nthin=nthin,
                                              a_scale=a, use_progress_meter=false)

# regenerate the consumed draws, in consumption order
Random.seed!(7)
nitw, half = niter ÷ nwalkers, nwalkers ÷ 2
partner = Int64[]; z = Float64[]; u = Float64[]
for t in 1:nitw, batch in 1:2
    ncos = batch == 1 ? (half+1:nwalkers) : (1:half)
    for nc in 1:half
        push!(partner, rand(ncos) - 1)                   # 0-based for the C-ABI
        push!(z, KissMCMC.cdf_g_inv(rand(), a))
        push!(u, rand())
    end
end

write(joinpath(outdir, "meta.txt"), "rosenbrock $nwalkers $nitw $(nburnin ÷ nwalkers) $nthin $a\n")
write(joinpath(outdir, "theta0s.f64"), reduce(hcat, theta0s))                       # d x nw
write(joinpath(outdir, "partner.i64"), partner)
write(joinpath(outdir, "z.f64"), z)
write(joinpath(outdir, "u.f64"), u)
ns = length(thetas[1])
chain = Array{Float64,3}(undef, 2, ns, nwalkers)
for w in 1:nwalkers, s in 1:ns
    chain[:, s, w] = thetas[w][s]
end
write(joinpath(outdir, "chain.f64"), chain)                                         # d x ns x nw
write(joinpath(outdir, "logp.f64"), reduce(hcat, logdensities))                     # ns x nw
write(joinpath(outdir, "accept_ratio.f64"), accept_ratio)
println("wrote reference trace and chains to $outdir")
