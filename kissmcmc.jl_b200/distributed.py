"""Multi-GPU drivers: one process per GPU, `torch.distributed` for the plumbing.

Two ways to use several B200s:

  emcee_independent   G separate ensembles, rank-distinct Philox key and walker-id range, NO
                      data-path collective; chains are gathered at the end only.
  emcee_sharded       ONE ensemble sharded by walker index: rank r updates positions
                      [r*S, (r+1)*S) of each half (S = nwalkers/2/G) and, after every half-step,
                      the updated half is all-gathered (in place) so that every rank can draw
                      partners from the whole complementary half (src/samplers.jl:250).  Draws are
                      keyed by the global walker index, so the result is bit-identical to the
                      single-GPU run of the same ensemble.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from ._lib import EXCHANGE_PUSH
from .api import LogDensity, Sampler, _require_plugin


def shard_range(nwalkers: int, rank: int, world: int):
    """(begin, count) of the positions of each half that `rank` updates."""
    nhalf = nwalkers // 2
    if nwalkers % 2 or nhalf % world:
        raise ValueError(f"nwalkers/2 = {nhalf} must be a multiple of the number of ranks ({world})")
    s = nhalf // world
    return rank * s, s


def assemble_shards(parts, nwalkers: int):
    """parts[r] = rank r's local arrays [2*S, ...] (its slice of half 0, then of half 1) ->
    the global walker order [nwalkers, ...]."""
    s = parts[0].shape[0] // 2
    return np.concatenate([p[:s] for p in parts] + [p[s:] for p in parts], axis=0)


class _DeviceArray:
    """Exposes a raw device pointer to torch through __cuda_array_interface__."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": tuple(shape), "typestr": typestr,
                                         "version": 2, "strides": None}


def x_tensor(sampler: Sampler) -> torch.Tensor:
    """The sampler's device-resident positions as a torch tensor [nw, d] (no copy)."""
    xptr, _, _ = sampler.device_ptrs()
    dev = torch.device("cuda", sampler.opts.device)
    return torch.as_tensor(_DeviceArray(xptr, (sampler.nw, sampler.d), "<f8"), device=dev)


def _gather_numpy(local: np.ndarray, group=None):
    world = dist.get_world_size(group)
    out = [None] * world
    dist.all_gather_object(out, local, group=group)
    return out


def emcee_sharded(logdensity, theta0s, *, niter, nburnin=None, nthin=1, a_scale=2.0, seed=0, group=None,
                  sampler_factory=None, x_view=None, exchange="allgather", push_opts=None):
    """emcee (src/samplers.jl:188-216) of one ensemble sharded over the ranks of `group`.

    Every rank passes the SAME theta0s ([nw] or [nw, d]).  Returns the reference 4-tuple for
    the WHOLE ensemble on every rank.  sampler_factory / x_view are test seams (CPU fakes).

    exchange="allgather": one launch per half-step + NCCL all-gather of the updated half.
    exchange="peer":      ONE persistent kernel per rank; partner rows are gathered straight from
                          the owner GPU's memory over NVLink (CUDA IPC) and the ranks meet at a
                          flag barrier in peer memory per half-step -- no collective at all.
    exchange="push":      ONE persistent kernel per rank, every rank holds ONLY its shard: the owners
                          of the passive half push the packed partner rows each peer's walkers will
                          ask for into the peer's receive ring (bulk stores over NVLink) and every
                          consumer starts as soon as ITS chunk's rows have landed -- no collective, no
                          cross-GPU barrier (csrc/kmc_push.cuh)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    th = np.asarray(theta0s, dtype=np.float64)
    scalar_theta = th.ndim == 1
    if nburnin is None:
        nburnin = niter // 2
    assert a_scale > 1
    nwalkers = len(th)
    assert nwalkers % 2 == 0, "Use an even number of walkers."
    niter_walker, nburnin_walker = niter // nwalkers, nburnin // nwalkers
    npar = 1 if scalar_theta else th.shape[1]
    assert nwalkers >= npar + 2, "Use more walkers: at least DOF+2, but better many more."
    begin, count = shard_range(nwalkers, rank, world)
    nhalf = nwalkers // 2

    import contextlib
    ctx = contextlib.nullcontext()
    if sampler_factory is None and exchange == "push":
        _require_plugin(logdensity)
        s = Sampler(logdensity, th, niter_walker, nburnin_walker, nthin, a_scale, seed, launch_mode=0,
                    shard=(begin, count), exchange=EXCHANGE_PUSH, **(push_opts or {}))
        try:
            allh = [None] * world
            dist.all_gather_object(allh, s.window_export(), group=group)
            s.window_attach(allh, rank)
            dist.barrier(group)                 # every rank's window is mapped and its initial state in place
            s.run(-1, sync=True)                # the kernels of all ranks synchronise among themselves
            dist.barrier(group)                 # nobody tears its window down while a peer still writes into it
            lth, llp, lar = s.results()
        finally:
            s.close()
        thetas = assemble_shards(_gather_numpy(lth, group), nwalkers)
        logp = assemble_shards(_gather_numpy(llp, group), nwalkers)
        ratio = assemble_shards(_gather_numpy(lar, group), nwalkers)
        if scalar_theta:
            thetas = thetas[:, :, 0]
        return thetas, ratio, logp, None
    if sampler_factory is None and exchange == "peer":
        _require_plugin(logdensity)
        s = Sampler(logdensity, th, niter_walker, nburnin_walker, nthin, a_scale, seed, launch_mode=0,
                    shard=(begin, count))
        try:
            hx, hf = s.ipc_export()
            allh = [None] * world
            dist.all_gather_object(allh, (hx, hf), group=group)
            s.set_peers([h[0] for h in allh], [h[1] for h in allh], rank)
            dist.barrier(group)                 # every rank's initial state is in place
            s.run(-1, sync=True)                # the kernels of all ranks synchronise among themselves
            dist.barrier(group)                 # nobody tears its memory down while a peer still reads it
            lth, llp, lar = s.results()
        finally:
            s.close()
        thetas = assemble_shards(_gather_numpy(lth, group), nwalkers)
        logp = assemble_shards(_gather_numpy(llp, group), nwalkers)
        ratio = assemble_shards(_gather_numpy(lar, group), nwalkers)
        if scalar_theta:
            thetas = thetas[:, :, 0]
        return thetas, ratio, logp, None
    if sampler_factory is None:
        _require_plugin(logdensity)
        s = Sampler(logdensity, th, niter_walker, nburnin_walker, nthin, a_scale, seed, launch_mode=1,
                    shard=(begin, count))
        # kernels and collectives are ordered through ONE explicit torch stream (a NULL handle would
        # mean "the sampler's own stream" to the library, so the legacy default stream is not used)
        st = torch.cuda.Stream(device=s.opts.device)
        s.set_stream(st.cuda_stream)
        ctx = torch.cuda.stream(st)
        xt = x_tensor(s)
    else:
        s = sampler_factory(logdensity, th, niter_walker, nburnin_walker, nthin, a_scale, seed, (begin, count))
        xt = x_view(s)
    try:
        with ctx:
            for h in range(2 * niter_walker):
                s.run_half(1)
                half = xt[(h & 1) * nhalf:((h & 1) + 1) * nhalf]          # the half that was just updated
                mine = half[begin:begin + count]
                if xt.is_cuda:
                    dist.all_gather_into_tensor(half.view(-1), mine.reshape(-1), group=group)   # in place (NCCL)
                else:
                    parts = [torch.empty_like(mine) for _ in range(world)]
                    dist.all_gather(parts, mine.clone(), group=group)
                    half.copy_(torch.cat(parts, dim=0))
        s.sync()
        lth, llp, lar = s.results()
    finally:
        s.close()
    thetas = assemble_shards(_gather_numpy(lth, group), nwalkers)
    logp = assemble_shards(_gather_numpy(llp, group), nwalkers)
    ratio = assemble_shards(_gather_numpy(lar, group), nwalkers)
    if scalar_theta:
        thetas = thetas[:, :, 0]
    return thetas, ratio, logp, None


def emcee_independent(logdensity, theta0s, *, niter, nburnin=None, nthin=1, a_scale=2.0, seed=0, group=None,
                      gather=True, sampler_factory=None):
    """G independent ensembles, one per rank, no communication during sampling.  Every rank
    passes its OWN theta0s.  With gather=True every rank returns the concatenation over ranks
    (rank-major: ensemble 0's walkers, then ensemble 1's, ...), ready for squash_walkers."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    th = np.asarray(theta0s, dtype=np.float64)
    scalar_theta = th.ndim == 1
    if nburnin is None:
        nburnin = niter // 2
    nwalkers = len(th)
    assert a_scale > 1 and nwalkers % 2 == 0, "Use an even number of walkers."
    niter_walker, nburnin_walker = niter // nwalkers, nburnin // nwalkers
    if sampler_factory is None:
        _require_plugin(logdensity)
        s = Sampler(logdensity, th, niter_walker, nburnin_walker, nthin, a_scale, seed,
                    walker_id_base=rank * nwalkers)           # disjoint walker ids => disjoint Philox streams
    else:
        s = sampler_factory(logdensity, th, niter_walker, nburnin_walker, nthin, a_scale, seed, rank * nwalkers)
    try:
        s.run(-1)
        lth, llp, lar = s.results()
    finally:
        s.close()
    if gather:
        lth = np.concatenate(_gather_numpy(lth, group), axis=0)
        llp = np.concatenate(_gather_numpy(llp, group), axis=0)
        lar = np.concatenate(_gather_numpy(lar, group), axis=0)
    if scalar_theta:
        lth = lth[:, :, 0]
    return lth, lar, llp, None
