"""Sharded-ensemble throughput (BASELINE.json configs[4]: 10-D Gaussian, 2^24 walkers in ONE ensemble).
    torchrun --nproc-per-node G profiles/sharded_bench.py [log2_nwalkers] [iters]
Each half-step: every rank updates its slice, then NCCL all-gathers the updated half in place."""
import json, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch, torch.distributed as dist
import bench, kissmcmc_b200 as km
from kissmcmc_b200 import distributed as kd

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 24
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
mode = sys.argv[3] if len(sys.argv) > 3 else "allgather"
wl = dict(bench.WORKLOADS["gaussian10d"], nw=1 << lg)
params, x0 = bench.make_inputs(wl, 1)
nw, d, nhalf = wl["nw"], wl["d"], wl["nw"] // 2
ld = km.LogDensity("gaussian", d, params, device=local)
begin, count = kd.shard_range(nw, rank, world)
if mode == "peer":
    s = km.Sampler(ld, x0, iters + 4, 0, 10**6, 2.0, 7, launch_mode=0, shard=(begin, count), device=local)
    if world > 1:
        hx, hf = s.ipc_export()
        allh = [None] * world
        dist.all_gather_object(allh, (hx, hf))
        s.set_peers([h[0] for h in allh], [h[1] for h in allh], rank)
        dist.barrier()
    elif os.environ.get("KMC_BULK1"):
        hx, hf = s.ipc_export()
        s.set_peers([hx], [hf], 0)
    s.run(4, sync=True)                       # warm-up: 4 iterations
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    s.run(iters, sync=True)
    ms_k, _ = s.last_run_ms()
    ms = torch.tensor([ms_k], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.barrier()
    if rank == 0:
        print(json.dumps({"mode": "sharded-peer", "n_gpus": world, "nwalkers": nw, "d": d, "iters": iters,
                          "ms_per_halfstep": ms.item() / (2 * iters), "walker_steps_per_s": nw * iters / (ms.item() * 1e-3)}))
    s.close()
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0)
s = km.Sampler(ld, x0, iters + 4, 0, 10**6, 2.0, 7, launch_mode=1, shard=(begin, count), device=local)
st = torch.cuda.Stream()
s.set_stream(st.cuda_stream)
xt = kd.x_tensor(s)
def halfsteps(n, h0):
    for h in range(h0, h0 + n):
        s.run_half(1)
        if world > 1:
            half = xt[(h & 1) * nhalf:((h & 1) + 1) * nhalf]
            dist.all_gather_into_tensor(half.view(-1), half[begin:begin + count].reshape(-1))
with torch.cuda.stream(st):
    halfsteps(8, 0)
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    halfsteps(2 * iters, 8)
    e1.record(st)
    torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"mode": "sharded-allgather", "n_gpus": world, "nwalkers": nw, "d": d, "iters": iters,
                      "ms_per_halfstep": ms.item() / (2 * iters), "walker_steps_per_s": nw * iters / (ms.item() * 1e-3),
                      "allgather_bytes_per_rank_per_halfstep": nhalf * d * 8}))
s.close()
if world > 1:
    dist.destroy_process_group()
