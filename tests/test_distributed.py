"""world_size-2 tests of the multi-GPU host logic on CPU (`gloo`), and on 2 GPUs (`-m gpu`, NCCL).

The CPU tests replace the CUDA sampler by a fake built from the oracle's primitives (draws,
log-density), so they exercise the real driver code: shard ranges, the per-half-step in-place
all-gather offsets, result assembly, independent-ensemble id ranges."""
import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import cases


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class FakeShardSampler:
    """CPU stand-in with the Sampler interface the drivers use; steps walkers with oracle primitives."""

    def __init__(self, orc, od, th, nitw, nbw, nthin, a, seed, shard=None, id_base=0):
        self.orc, self.od = orc, od
        self.x = torch.from_numpy(np.array(th, dtype=np.float64).reshape(len(th), -1))
        self.nw, self.d = self.x.shape
        self.nhalf = self.nw // 2
        self.begin, self.count = shard if shard else (0, self.nhalf)
        self.lp = od.eval(self.x.numpy())
        self.nitw, self.nbw, self.nthin, self.a, self.seed, self.id_base = nitw, nbw, nthin, a, seed, id_base
        self.ns = (nitw - nbw) // nthin
        self.nl = 2 * self.count
        self.h = 0
        self.nacc = np.zeros(self.nw, dtype=np.int64)
        self.chain_x = np.zeros((self.nl, self.ns, self.d))
        self.chain_lp = np.zeros((self.nl, self.ns))
        self.sia = math.sqrt(1.0 / a)
        self.span = math.sqrt(a) - self.sia

    def run_half(self, n=1):
        for _ in range(n):
            t, batch = self.h >> 1, self.h & 1
            nref = t + 1 - self.nbw
            a0, p0 = (self.nhalf, 0) if batch else (0, self.nhalf)
            x = self.x.numpy()
            for i in range(self.begin, self.begin + self.count):
                k = a0 + i
                pl, uz, u = self.orc.draw(self.seed, self.id_base + k, t, batch, self.nhalf)
                s = uz * self.span + self.sia
                z = s * s
                y = x[p0 + pl] + z * (x[k] - x[p0 + pl])
                p1 = self.od.logpdf(y)
                lu = math.log(u) if u > 0 else -math.inf
                if ((self.d - 1) * math.log(z) + p1) - self.lp[k] >= lu:
                    x[k] = y
                    self.lp[k] = p1
                    self.nacc[k] += 1
                if nref > 0 and nref % self.nthin == 0:
                    row = (self.count if batch else 0) + (i - self.begin)
                    self.chain_x[row, nref // self.nthin - 1] = x[k]
                    self.chain_lp[row, nref // self.nthin - 1] = self.lp[k]
            if batch == 1 and nref == 0:
                self.nacc[:] = 0
            self.h += 1

    def run(self, n=-1):
        self.run_half(2 * self.nitw - self.h)

    def sync(self):
        pass

    def close(self):
        pass

    def results(self):
        idx = np.r_[self.begin:self.begin + self.count, self.nhalf + self.begin:self.nhalf + self.begin + self.count]
        return self.chain_x, self.chain_lp, self.nacc[idx] / (self.nitw - self.nbw)


def _worker(rank, world, port, mode, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import kissmcmc_b200 as km
    from oracle import oracle as orc
    name, d, params, th0, rad = cases.plugin_specs()["rosenbrock"]
    od = orc.Density(name, d, params)
    nw, nitw, nbw, nthin, a, seed = 16, 6, 2, 2, 2.0, 5
    try:
        if mode == "sharded":
            x0 = cases.ball(th0, rad, nw, 1)
            fac = lambda ld, th, ni, nb, nt, aa, sd, shard: FakeShardSampler(orc, od, th, ni, nb, nt, aa, sd, shard=shard)
            out = km.distributed.emcee_sharded(None, x0, niter=nitw * nw, nburnin=nbw * nw, nthin=nthin, a_scale=a,
                                               seed=seed, sampler_factory=fac, x_view=lambda s: s.x)
        else:
            x0 = cases.ball(th0, rad, nw, 10 + rank)
            fac = lambda ld, th, ni, nb, nt, aa, sd, idb: FakeShardSampler(orc, od, th, ni, nb, nt, aa, sd, id_base=idb)
            out = km.distributed.emcee_independent(None, x0, niter=nitw * nw, nburnin=nbw * nw, nthin=nthin,
                                                   a_scale=a, seed=seed, sampler_factory=fac)
        q.put((rank, out[0], out[1], out[2]))
    finally:
        dist.destroy_process_group()


def _run(mode, world=2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, mode, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    [p.join(30) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    return res


def test_shard_range_and_assembly(km):
    assert km.distributed.shard_range(16, 1, 2) == (4, 4)
    with pytest.raises(ValueError):
        km.distributed.shard_range(20, 0, 4)
    a = np.arange(16).reshape(16, 1)
    parts = [np.concatenate([a[r * 2:(r + 1) * 2], a[8 + r * 2:8 + (r + 1) * 2]]) for r in range(4)]
    assert np.array_equal(km.distributed.assemble_shards(parts, 16), a)


def test_sharded_two_ranks_equals_single_process_oracle(orc):
    """2 ranks x half of each half == the oracle's whole-ensemble run, bit for bit, on every rank."""
    res = _run("sharded")
    name, d, params, th0, rad = cases.plugin_specs()["rosenbrock"]
    want = orc.emcee(orc.Density(name, d, params), cases.ball(th0, rad, 16, 1), 6, 2, 2, 2.0, seed=5)
    for rank, th, ar, lp in res:
        assert np.array_equal(th, want["chain_x"]) and np.array_equal(lp, want["chain_lp"])
        assert np.array_equal(ar, want["accept_ratio"])


def test_independent_two_ranks(orc):
    """Each rank's ensemble equals the oracle run with that rank's walker-id base; gather is rank-major."""
    res = _run("independent")
    name, d, params, th0, rad = cases.plugin_specs()["rosenbrock"]
    od = orc.Density(name, d, params)
    wants = []
    for r in range(2):
        f = FakeShardSampler(orc, od, cases.ball(th0, rad, 16, 10 + r), 6, 2, 2, 2.0, 5, id_base=r * 16)
        f.run()
        wants.append(f.results())
    for rank, th, ar, lp in res:
        assert th.shape == (32, 2, 2)
        assert np.array_equal(th, np.concatenate([w[0] for w in wants]))
        assert np.array_equal(ar, np.concatenate([w[2] for w in wants]))
    assert not np.array_equal(res[0][1][:16], res[0][1][16:])     # the two ensembles differ


# ---------------------------------------------------------------------------------- GPU (NCCL)

def _gpu_worker(rank, world, port, q, exchange="allgather"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import kissmcmc_b200 as km
    try:
        ld = km.LogDensity("gaussian", 10, cases.plugin_specs()["mvn10"][2], device=rank)
        x0 = cases.ball(np.zeros(10), 0.1, 4096, 3)
        out = km.distributed.emcee_sharded(ld, x0, niter=30 * 4096, nburnin=10 * 4096, nthin=5, seed=11,
                                           exchange=exchange)
        q.put((rank, out[0], out[1], out[2]))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("exchange", ["allgather", "peer"])
def test_sharded_two_gpus_equals_one_gpu(km, exchange):
    """One ensemble over 2 GPUs == 1 GPU, bit for bit: with an NCCL all-gather of the updated half per
    half-step, and with the fused peer mode (partner rows gathered over NVLink inside one persistent
    kernel per rank, flag barrier in peer memory)."""
    if km.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gpu_worker, args=(r, 2, port, q, exchange)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=300) for _ in range(2)], key=lambda t: t[0])
    [p.join(60) for p in procs]
    ld = km.LogDensity("gaussian", 10, cases.plugin_specs()["mvn10"][2])
    x0 = cases.ball(np.zeros(10), 0.1, 4096, 3)
    th, ar, lp, _ = km.emcee(ld, x0, niter=30 * 4096, nburnin=10 * 4096, nthin=5, seed=11, use_progress_meter=False)
    for rank, sth, sar, slp in res:
        assert np.array_equal(sth, th) and np.array_equal(slp, lp) and np.array_equal(sar, ar)


@pytest.mark.gpu
@pytest.mark.parametrize("case,nw", [("mvn10", 4096), ("rosenbrock", 2048), ("exponential3", 600)])
def test_two_shards_on_one_gpu_equal_unsharded(km, case, nw):
    """The sharded-ensemble logic exercised on ONE GPU: two samplers own the two halves of each
    half-ensemble; after every half-step the updated slices are exchanged by device copies (what the
    all-gather does across GPUs).  Global walker ids key the Philox draws, so the result is bit-identical
    to the unsharded run."""
    name, d, params, th0, rad = cases.plugin_specs()[case]
    ld = km.LogDensity(name, d, params)
    x0 = cases.ball(th0, rad, nw, 3)
    if name == "exponential":
        x0 = np.abs(x0)
    nitw, nbw, nthin = 12, 4, 2
    th, ar, lp, _ = km.emcee(ld, x0, niter=nitw * nw, nburnin=nbw * nw, nthin=nthin, seed=19, use_progress_meter=False)
    nhalf, S = nw // 2, nw // 4
    ss = [km.Sampler(ld, x0, nitw, nbw, nthin, 2.0, 19, launch_mode=1, shard=km.distributed.shard_range(nw, r, 2))
          for r in range(2)]
    xs = [km.distributed.x_tensor(s) for s in ss]
    for h in range(2 * nitw):
        for s in ss:
            s.run_half(1, sync=True)
        lo = (h & 1) * nhalf
        xs[0][lo + S:lo + 2 * S].copy_(xs[1][lo + S:lo + 2 * S])      # rank 1's slice -> rank 0
        xs[1][lo:lo + S].copy_(xs[0][lo:lo + S])                      # rank 0's slice -> rank 1
        torch.cuda.synchronize()
    parts = [s.results() for s in ss]
    [s.close() for s in ss]
    got_th = km.distributed.assemble_shards([p[0] for p in parts], nw)
    got_lp = km.distributed.assemble_shards([p[1] for p in parts], nw)
    got_ar = km.distributed.assemble_shards([p[2] for p in parts], nw)
    assert np.array_equal(got_th, th) and np.array_equal(got_lp, lp) and np.array_equal(got_ar, ar)


@pytest.mark.gpu
def test_single_rank_peer_mode_equals_plain(km):
    """Peer mode with one rank runs the cross-GPU code path alone (bulk-copy gathers through the
    peer pointer table, flag barrier on its own flags): bit-identical to the plain run."""
    name, d, params, th0, rad = cases.plugin_specs()["mvn10"]
    ld = km.LogDensity(name, d, params)
    x0 = cases.ball(th0, rad, 20000, 3)
    out = []
    for peer in (False, True):
        s = km.Sampler(ld, x0, 10, 4, 2, 2.0, 5, shard=(0, 10000))
        if peer:
            hx, hf = s.ipc_export()
            s.set_peers([hx], [hf], 0)
        s.run(3)
        s.run(-1)
        out.append(s.results() + s.state())
        s.close()
    for a, b in zip(*out):
        assert np.array_equal(a, b)
