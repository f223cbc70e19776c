# KissMCMC.CUDABackend -- thin `ccall` layer over libkissmcmc_cuda.so (include/kissmcmc_cuda.h).
#
# Drop this file next to src/samplers.jl of KissMCMC.jl and add `include("CUDABackend.jl")` to
# src/KissMCMC.jl (after line 11).  It keeps the reference's call shapes
#
#     emcee(logdensity, theta0s; niter, nburnin, nthin, a_scale, use_progress_meter, hasblob, ...)
#     make_theta0s(theta0, ball_radius, logdensity, nwalkers; ...)
#     squash_walkers(thetas, accept_ratio, logdensities, blobs; ...)   (unchanged: host side)
#
# with ONE difference: `logdensity` is a `LogDensity` plugin descriptor instead of a closure
# (closures cannot cross the C-ABI onto the GPU; there is no CPU fallback and no CUDA.jl).
#
# NOTE: Julia is not installed in the build image, so this module has not been executed
# there; the identical C-ABI is exercised by the Python ctypes harness (kissmcmc.jl_b200/api.py)
# and the GPU parity tests.
module CUDABackend

export LogDensity, exponential, rosenbrock, gaussian, lognormal, emcee, make_theta0s

using LinearAlgebra: cholesky, Symmetric, diag, inv
import ..KissMCMC: squash_walkers   # host-side, reused as is (src/samplers.jl:372-428)

const LIB = Ref{String}(get(ENV, "KISSMCMC_CUDA_LIB", "libkissmcmc_cuda"))

const MODE_PHILOX = Int32(0)
const MODE_REPLAY = Int32(1)

struct KmcError <: Exception
    code::Int32
    msg::String
end
Base.showerror(io::IO, e::KmcError) = print(io, "libkissmcmc_cuda error $(e.code): $(e.msg)")

last_error() = unsafe_string(ccall((:kmc_last_error, LIB[]), Cstring, ()))
check(rc::Int32) = rc == 0 ? nothing : throw(KmcError(rc, last_error()))

# mirrors `struct kmc_emcee_opts` field for field
struct EmceeOpts
    niter_walker::Int64
    nburnin_walker::Int64
    nthin::Int64
    a_scale::Float64
    seed::UInt64
    mode::Int32
    device::Int32
    walker_id_base::Int64
    launch_mode::Int32
    exchange::Int32
    shard_begin::Int64
    shard_count::Int64
    push_chunk::Int32
    push_cap::Int32
    push_lag::Int32
    reserved::Int32
end

# ------------------------------------------------------------------ log-density plugins
"A device log-density plugin instance (replaces the closure `pdf`, src/samplers.jl:257)."
mutable struct LogDensity
    handle::Ptr{Cvoid}
    name::String
    d::Int
    function LogDensity(name::AbstractString, d::Integer, params::Vector{Float64}=Float64[];
                        data::Union{Nothing,Array}=nothing, device::Integer=0)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        dptr = data === nothing ? C_NULL : pointer(data)
        dbytes = data === nothing ? 0 : sizeof(data)
        GC.@preserve params data check(ccall((:kmc_density_create, LIB[]), Int32,
            (Cstring, Int32, Ptr{Float64}, Int64, Ptr{Cvoid}, Int64, Int32, Ref{Ptr{Cvoid}}),
            name, d, params, length(params), dptr, dbytes, device, h))
        obj = new(h[], String(name), Int(d))
        finalizer(o -> ccall((:kmc_density_destroy, LIB[]), Int32, (Ptr{Cvoid},), o.handle), obj)
        return obj
    end
end

"README.md:15  `logpdf(x) = x<0 ? -Inf : -x`"
exponential(d::Integer=1; device=0) = LogDensity("exponential", d; device=device)
"test/runtests.jl:68  `-(b*(x[2]-x[1]^2)^2 + (a-x[1])^2)/temper`"
rosenbrock(a=1.0, b=100.0, temper=20.0; device=0) = LogDensity("rosenbrock", 2, Float64[a, b, temper]; device=device)
"MvNormal(mean, cov) (test/runtests.jl:53,61): params = [mu; vec(A'); lognorm], A = chol(inv(cov))' so that |A(x-mu)|^2 is the Mahalanobis form"
function gaussian(mean, cov; device=0)
    mu = Float64.(vcat(mean))
    d = length(mu)
    prec = inv(reshape(Float64.(collect(cov)), d, d))
    Lc = cholesky(Symmetric((prec + prec') / 2)).L          # prec = L L'
    A = Matrix(Lc')                                          # y = A (x - mu)
    lognorm = sum(log.(diag(Lc))) - 0.5 * d * log(2pi)
    params = vcat(mu, vec(permutedims(A)), lognorm)          # A row-major
    return LogDensity("gaussian", d, params; device=device)
end
"LogNormal(mu, sigma) (test/runtests.jl:56)"
lognormal(mu=0.0, sigma=1.0; device=0) =
    LogDensity("lognormal", 1, Float64[mu, sigma, log(sigma) + 0.5 * log(2pi)]; device=device)

"Batched evaluation: `thetas` is d x n (column = one point); returns n log-densities."
function (ld::LogDensity)(thetas::AbstractMatrix{Float64})
    @assert size(thetas, 1) == ld.d
    th = Matrix{Float64}(thetas)
    out = Vector{Float64}(undef, size(th, 2))
    GC.@preserve th out check(ccall((:kmc_density_eval, LIB[]), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}), ld.handle, th, size(th, 2), out))
    return out
end
(ld::LogDensity)(theta::AbstractVector{<:Real}) = ld(reshape(Float64.(theta), ld.d, 1))[1]
(ld::LogDensity)(theta::Real) = ld(reshape(Float64[theta], 1, 1))[1]

# ------------------------------------------------------------------ emcee (src/samplers.jl:188-216)
to_matrix(theta0s::AbstractVector{<:Real}) = reshape(Float64.(theta0s), 1, :)
to_matrix(theta0s::AbstractVector) = Float64.(reduce(hcat, theta0s))      # d x nw, column-major

"""
    emcee(logdensity::LogDensity, theta0s; niter=10^5, nburnin=niter÷2, nthin=1, a_scale=2.0,
          use_progress_meter=true, hasblob=false, seed=0, replay=nothing)

Same semantics and 4-tuple as `KissMCMC.emcee`: `(thetas, accept_ratio, logdensities, nothing)`
with `thetas::Vector{Vector{T}}` (one chain per walker).  `replay=(partner, z, u)` uploads the
draws (partner = 1-based walker index as in Julia; converted to 0-based for the ABI).
"""
function emcee(ld::LogDensity, theta0s; niter=10^5, nburnin=niter ÷ 2, nthin=1, a_scale=2.0,
               use_progress_meter=true, hasblob=false, init_blobs=nothing, reduce_blob! =nothing,
               seed::Integer=0, replay=nothing, device::Integer=0)
    hasblob && error("hasblob=true is not supported by the CUDA backend (blobs are host objects)")
    x0 = to_matrix(deepcopy(theta0s))                                       # :198
    @assert a_scale > 1                                                     # :200
    nwalkers = size(x0, 2)
    @assert iseven(nwalkers) "Use an even number of walkers."               # :202
    niter_walker = niter ÷ nwalkers                                         # :203
    nburnin_walker = nburnin ÷ nwalkers                                     # :204
    d = size(x0, 1)
    @assert nwalkers >= d + 2 "Use more walkers: at least DOF+2, but better many more."   # :205

    opts = Ref(EmceeOpts(niter_walker, nburnin_walker, nthin, a_scale, UInt64(seed),
                         replay === nothing ? MODE_PHILOX : MODE_REPLAY, device, 0, 0, 0, 0, 0, 0, 0, 0, 0))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve x0 check(ccall((:kmc_emcee_create, LIB[]), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Int32, Ref{EmceeOpts}, Ref{Ptr{Cvoid}}),
        ld.handle, x0, nwalkers, d, opts, h))
    s = h[]
    try
        if replay !== nothing
            partner = Int64.(replay[1]) .- 1
            z, u = Float64.(replay[2]), Float64.(replay[3])
            GC.@preserve partner z u check(ccall((:kmc_emcee_set_replay, LIB[]), Int32,
                (Ptr{Cvoid}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Int64),
                s, partner, z, u, length(partner) ÷ nwalkers))
        end
        if use_progress_meter && niter_walker > 0          # coarse: at most 20 polls, never per iteration
            done = 0
            for c in 1:min(20, niter_walker)
                upto = (niter_walker * c) ÷ min(20, niter_walker)
                check(ccall((:kmc_emcee_run, LIB[]), Int32, (Ptr{Cvoid}, Int64), s, upto - done))
                done = upto
                it, m, sd, outl = Ref{Int64}(0), Ref{Float64}(0), Ref{Float64}(0), Ref{Int64}(0)
                check(ccall((:kmc_emcee_progress, LIB[]), Int32,
                    (Ptr{Cvoid}, Ref{Int64}, Ref{Float64}, Ref{Float64}, Ref{Int64}), s, it, m, sd, outl))
                nn = max(1, it[] > nburnin_walker ? it[] - nburnin_walker : it[])
                print(stderr, "\remcee, niter=$niter, nwalkers=$nwalkers: $(100done ÷ niter_walker)%  ",
                      "accept_ratio_mean=$(round(m[]/nn, sigdigits=3)) accept_ratio_std=$(round(sd[]/nn, sigdigits=3)) ",
                      "accept_ratio_outliers=$(outl[]) burnin_phase=$(it[] <= nburnin_walker)")
            end
            println(stderr)
        else
            check(ccall((:kmc_emcee_run, LIB[]), Int32, (Ptr{Cvoid}, Int64), s, -1))
        end
        nsr = Ref{Int64}(0)
        check(ccall((:kmc_emcee_nsamples, LIB[]), Int32, (Ptr{Cvoid}, Ref{Int64}), s, nsr))
        ns = nsr[]
        th = Array{Float64,3}(undef, d, ns, nwalkers)       # (ntheta, nsamples, nchains), src/analysis.jl:143
        lp = Matrix{Float64}(undef, ns, nwalkers)
        ar = Vector{Float64}(undef, nwalkers)
        GC.@preserve th lp ar check(ccall((:kmc_emcee_copy_results, LIB[]), Int32,
            (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), s, th, lp, ar))
        scalar = eltype(theta0s) <: Real
        thetas = scalar ? [th[1, :, w] for w in 1:nwalkers] :
                          [[th[:, i, w] for i in 1:ns] for w in 1:nwalkers]
        logdensities = [lp[:, w] for w in 1:nwalkers]
        return thetas, ar, logdensities, nothing            # blobs: nothing (:221, test/emcee.jl:33)
    finally
        ccall((:kmc_emcee_destroy, LIB[]), Int32, (Ptr{Cvoid},), s)
    end
end

# ------------------------------------------------------------------ make_theta0s (src/samplers.jl:311-349)
"""
    make_theta0s(theta0, ball_radius, logdensity::LogDensity, nwalkers; ball_radius_halfing_steps=7, ntries=100)

Reference loop semantics (including the cumulative radius halving of :326) with the density
calls batched through the plugin: all pending walkers are tried at once.
"""
function make_theta0s(theta0::T, ball_radius, ld::LogDensity, nwalkers;
                      ball_radius_halfing_steps=7, ntries=100, hasblob=false) where T
    hasblob && error("hasblob=true is not supported by the CUDA backend")
    npara = length(theta0)
    br = ball_radius isa Number ? ones(npara) * ball_radius : Float64.(collect(ball_radius))   # :316-318
    @assert length(br) == npara                                                                # :319
    th0 = Float64.(vcat(theta0))
    out = Matrix{Float64}(undef, npara, nwalkers)
    found = falses(nwalkers)
    i0 = 1
    while i0 <= nwalkers
        pend = collect(i0:nwalkers)
        for j in 1:ntries                                                   # k = 1: radius factor 1
            tmp = th0 .+ randn(npara, length(pend)) .* br
            ok = ld(tmp) .> -Inf                                            # :338
            out[:, pend[ok]] = tmp[:, ok]
            found[pend[ok]] .= true
            pend = pend[.!ok]
            isempty(pend) && break
        end
        isempty(pend) && break
        f = pend[1]                        # first walker whose k=1 tries all failed: redo the later ones
        found[f+1:end] .= false
        for k in 2:ball_radius_halfing_steps                                # :324
            br = br .* (1 / 2^(k - 1))                                      # :326 cumulative, never reset
            for j in 1:ntries
                tmp = th0 .+ randn(npara, 1) .* br
                if ld(tmp)[1] > -Inf
                    out[:, f] = tmp; found[f] = true
                    break
                end
            end
            found[f] && break
        end
        i0 = f + 1
    end
    cols = findall(found)
    return T <: Number ? [out[1, c] for c in cols] : [T(out[:, c]) for c in cols]
end

end # module
