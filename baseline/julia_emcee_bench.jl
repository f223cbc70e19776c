# julia_emcee_bench.jl -- NOT RUN in the build image (no Julia there).  Times the UNMODIFIED
# reference KissMCMC.emcee (threaded over the active half, src/samplers.jl:248) on the bench.py
# workload so that anyone with Julia can fill the "Julia threaded CPU" column:
#
#   JULIA_NUM_THREADS=$(nproc) julia --project=/path/to/KissMCMC.jl baseline/julia_emcee_bench.jl [nwalkers] [iters]
using KissMCMC, Random

nwalkers = length(ARGS) >= 1 ? parse(Int, ARGS[1]) : 2^20
iters = length(ARGS) >= 2 ? parse(Int, ARGS[2]) : 50                 # iterations per walker (bounded sample)
rosen(x) = -(100 * (x[2] - x[1]^2)^2 + (1 - x[1])^2) / 20            # test/runtests.jl:68
theta0s = [0.1 .* randn(2) for _ in 1:nwalkers]
emcee(rosen, theta0s; niter=nwalkers * 2, use_progress_meter=false)  # compile
t = @elapsed emcee(rosen, theta0s; niter=nwalkers * iters, nthin=max(1, iters ÷ 5), use_progress_meter=false)
println("{\"impl\": \"julia-reference\", \"threads\": $(Threads.nthreads()), \"walker_steps_per_s\": $(nwalkers * iters / t), ",
        "\"nwalkers\": $nwalkers, \"iters\": $iters, \"seconds\": $t}")
