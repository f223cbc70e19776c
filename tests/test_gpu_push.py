"""GPU tests (`-m gpu`) of the sharded ensemble with owner-computes pushes (csrc/kmc_push.cuh) and of the
library-owned multi-GPU entry points (kmc_emcee_create_multi / kmc_multi_*), through the C-ABI.

The bar is the one of SURVEY.md section 8e: ONE ensemble sharded over G ranks gives the chains of the single-GPU run
of the same ensemble BIT FOR BIT (and therefore the oracle's, which the single-GPU path is pinned to).  The parallel
region replaced is /root/reference/src/samplers.jl:246-273.

Most cases run G sub-samplers on ONE GPU (a device ordinal repeated in `devices`): the same kernel, flags, receive
ring and bulk stores, with the peers' windows being ordinary device memory -- so the driver's 1-GPU box covers the
whole protocol.  The 2-GPU cases (NVLink peer memory: one process, and one process per GPU over CUDA IPC) skip on a
1-GPU box."""
import os
import socket

import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu


def _case(km, case, nw, seed=3):
    name, d, params, th0, rad = cases.plugin_specs()[case]
    x0 = cases.ball(th0, rad, nw, seed)
    if name == "exponential":
        x0 = np.abs(x0)
    return km.LogDensity(name, d, params), x0


def _single(km, ld, x0, nitw, nbw, nthin, a=2.0, seed=11):
    s = km.Sampler(ld, x0, nitw, nbw, nthin, a, seed)
    s.run(-1)
    out = s.results()
    s.close()
    return out


def _multi(km, ld, x0, nitw, nbw, nthin, devices, a=2.0, seed=11, chunks=None, **kw):
    m = km.MultiSampler(ld, x0, nitw, nbw, nthin, a, seed, devices=devices, sharded=True, **kw)
    for c in chunks or ():
        m.run(c)
    m.run(-1)
    out = m.results()
    ms = m.last_run_ms()
    m.close()
    assert ms > 0
    return out


def _same(a, b):
    for u, v, what in zip(a, b, ("chains", "log-densities", "accept ratios")):
        assert u.shape == v.shape, what
        assert np.array_equal(u, v), what


@pytest.mark.parametrize("case,nw,G,kw", [
    ("mvn10", 4096, 2, {}),                                   # BASELINE.json configs[4]'s density, 2 ranks
    ("mvn10", 8192, 4, {}),                                   # 4 ranks, chunk 512: two rounds per chunk
    ("mvn10", 16384, 8, {}),                                  # 8 ranks, chunk 1024: four rounds per chunk
    ("mvn10", 4096, 2, dict(push_cap=8)),                     # tiny ring slots: most rows take the owner-read fallback
    ("mvn10", 6000, 2, dict(push_chunk=384, push_lag=1)),     # ragged: S = 1500 is not a multiple of the chunk
    ("mvn10", 6000, 3, dict(push_chunk=100, push_lag=3)),     # 3 ranks, partial warps in every chunk
    ("mvn10", 8192, 4, dict(push_lag=-1)),                    # the adaptive task hand-out (two counters)
    ("mvn10", 6000, 2, dict(push_chunk=384, push_lag=-1)),    # adaptive, ragged
    ("rosenbrock", 2048, 2, {}),                              # d = 2 rows (16 bytes)
    ("mvn2", 1024, 4, dict(push_chunk=64)),
])
def test_sharded_on_one_gpu_equals_single_sampler(km, case, nw, G, kw):
    ld, x0 = _case(km, case, nw)
    nitw, nbw, nthin = 14, 4, 3
    want = _single(km, ld, x0, nitw, nbw, nthin)
    got = _multi(km, ld, x0, nitw, nbw, nthin, [0] * G, **kw)
    _same(got, want)


def test_sharded_in_several_launches_and_burnin_reset(km):
    """Chunked runs (several cooperative launches, flag epochs and ring parities carried across them), burn-in
    ending inside a launch (counter reset :285-288), nthin > 1."""
    ld, x0 = _case(km, "mvn10", 4096)
    want = _single(km, ld, x0, 21, 7, 2)
    got = _multi(km, ld, x0, 21, 7, 2, [0, 0], chunks=[1, 5, 1, 3])
    _same(got, want)
    assert want[2].min() >= 0 and want[2].max() <= 1


def test_sharded_matches_the_oracle(km, orc):
    name, d, params, th0, rad = cases.plugin_specs()["mvn10"]
    x0 = cases.ball(th0, rad, 1024, 5)
    want = orc.emcee(orc.Density(name, d, params), x0, 12, 4, 2, 2.0, seed=77, nthreads=4)
    got = _multi(km, km.LogDensity(name, d, params), x0, 12, 4, 2, [0, 0], seed=77)
    assert np.array_equal(got[0], want["chain_x"])
    assert np.array_equal(got[1], want["chain_lp"])
    assert np.array_equal(got[2], want["accept_ratio"])


def test_push_exchange_with_one_rank_equals_plain(km):
    """G = 1: the push kernel's dynamic task loop alone (no peers)."""
    ld, x0 = _case(km, "mvn10", 20000)
    want = _single(km, ld, x0, 9, 3, 2)
    s = km.Sampler(ld, x0, 9, 3, 2, 2.0, 11, shard=(0, 10000), exchange=km.EXCHANGE_PUSH)
    s.run(4)
    s.run(-1)
    got = s.results()
    x, lp, na = s.state()
    s.close()
    _same(got, want)
    assert x.shape == (20000, 10) and np.array_equal(lp, ld.eval(x))


def test_push_refuses_what_it_cannot_do(km):
    ld, x0 = _case(km, "mvn10", 4096)
    with pytest.raises(km.KmcError):      # unequal shards
        km.Sampler(ld, x0, 4, 0, 1, 2.0, 1, shard=(0, 1000), exchange=km.EXCHANGE_PUSH)
    s = km.Sampler(ld, x0, 4, 0, 1, 2.0, 1, shard=(0, 1024), exchange=km.EXCHANGE_PUSH)
    with pytest.raises(km.KmcError):      # two ranks, windows not attached
        s.run(-1)
    s.close()
    ld3, x3 = _case(km, "exponential3", 600)
    with pytest.raises(km.KmcError):      # odd d: rows are not 16-byte multiples
        km.Sampler(ld3, x3, 4, 0, 1, 2.0, 1, shard=(0, 150), exchange=km.EXCHANGE_PUSH)


def test_emcee_devices_keyword(km):
    ld, x0 = _case(km, "rosenbrock", 1024)
    want = km.emcee(ld, x0, niter=20 * 1024, nthin=2, seed=4, use_progress_meter=False)
    got = km.emcee(ld, x0, niter=20 * 1024, nthin=2, seed=4, use_progress_meter=False, devices=[0, 0])
    for u, v in zip(got[:3], want[:3]):
        assert np.array_equal(u, v)
    assert got[3] is None


def test_independent_ensembles_through_the_library(km):
    ld, _ = _case(km, "rosenbrock", 512)
    x0 = np.stack([cases.ball([0.0, 0.0], 0.1, 512, 40 + r) for r in range(3)])
    th, ar, lp, _ = km.emcee(ld, x0, niter=16 * 512, nthin=2, seed=9, use_progress_meter=False, devices=[0, 0, 0],
                             sharded=False)
    assert th.shape == (3 * 512, 4, 2) and ar.shape == (3 * 512,)
    for r in range(3):
        s = km.Sampler(ld, x0[r], 16, 8, 2, 2.0, 9, walker_id_base=r * 512)
        s.run(-1)
        wth, wlp, war = s.results()
        s.close()
        sl = slice(r * 512, (r + 1) * 512)
        assert np.array_equal(th[sl], wth) and np.array_equal(lp[sl], wlp) and np.array_equal(ar[sl], war)


# ---------------------------------------------------------------------------------- 2 real GPUs

def test_sharded_two_gpus_one_process(km):
    """NVLink peer memory inside one process (cudaDeviceEnablePeerAccess): devices [0, 1]."""
    if km.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ld, x0 = _case(km, "mvn10", 1 << 16)
    want = _single(km, ld, x0, 12, 4, 2)
    got = _multi(km, ld, x0, 12, 4, 2, [0, 1])
    _same(got, want)
    got = _multi(km, ld, x0, 12, 4, 2, [0, 1], push_cap=64)   # with owner-read fallbacks over NVLink
    _same(got, want)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _ipc_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import kissmcmc_b200 as km
    try:
        ld = km.LogDensity("gaussian", 10, cases.plugin_specs()["mvn10"][2], device=rank)
        x0 = cases.ball(np.zeros(10), 0.1, 1 << 16, 3)
        out = km.distributed.emcee_sharded(ld, x0, niter=12 << 16, nburnin=4 << 16, nthin=2, seed=11, exchange="push")
        q.put((rank, out[0], out[1], out[2]))
    finally:
        dist.destroy_process_group()


def test_sharded_two_gpus_one_process_per_gpu(km):
    """The torchrun shape: one process per GPU, windows exchanged as CUDA IPC handles."""
    if km.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ipc_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=300) for _ in range(2)], key=lambda t: t[0])
    [p.join(60) for p in procs]
    ld, x0 = _case(km, "mvn10", 1 << 16)
    th, lp, ar = _single(km, ld, x0, 12, 4, 2)
    for rank, sth, sar, slp in res:
        assert np.array_equal(sth, th) and np.array_equal(slp, lp) and np.array_equal(sar, ar)
