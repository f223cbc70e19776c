"""CPU model of the push exchange's protocol (csrc/kmc_push.cuh): the packing order of a message, the receiver's row
arithmetic, the task order and its deadlock-freedom, the ring / flag indexing.  numpy + plain Python, no GPU: the GPU
tests (tests/test_gpu_push.py) say whether the kernel implements this model; these say the model itself is sound for
every shape the host code can pick (ragged shards, partial warps, slots that overflow their capacity)."""
import re
from pathlib import Path

import numpy as np
import pytest

CSRC = Path(__file__).resolve().parent.parent / "kissmcmc.jl_b200" / "csrc"
SRC = (CSRC / "kmc_push.cuh").read_text()
T = 256          # kPushThreads
WARPS = T // 32


def test_constants_the_model_assumes():
    """The model below is written for the build default: CTA-wide tasks of 256 threads, chunks of up to 4 rounds."""
    assert re.search(r"#define KMC_PUSH_THREADS (\d+)", SRC).group(1) == str(T)
    assert "kPushMaxRounds = kPushThreads == 32 ? 8 : 1024 / kPushThreads" in SRC      # 4 rounds of 256 = 1024 walkers
    assert "kPushHeader = 4 * kPushMaxRounds" in SRC                                   # one u32 per round
    assert re.search(r"constexpr int kPushMaxRanks = (\d+)", SRC).group(1) == "8"
    max_cap = int(re.search(r"kPushMaxCap = kPushThreads == 32 \? (\d+) : (\d+)", SRC).group(2))
    assert max_cap == 384 and 2 * T * 16 * 8 + 16 + max_cap * 16 * 8 < 227 * 1024     # d = 16 fits one CTA's shared memory


def sender_pack(owner_of, me, chunk, cap):
    """push(c, dest) on owner `me`: the hits among the chunk's walkers, packed in (round, warp, lane) = walker order;
    returns (header[rounds], rows sent = list of walker offsets, rows that did not fit)."""
    rounds = (chunk + T - 1) // T
    hits = [off for off in range(chunk) if owner_of[off] == me]
    header = [sum(1 for off in hits if off < g * T) for g in range(rounds)]     # the prefix at (round g, warp 0)
    return header, hits[:cap], hits[cap:]


def receiver_row(owner_of, off, header):
    """update(c, g) on the receiver: walker `off`'s row in its owner's message = header[g] + rank among the ROUND's
    walkers with the same owner (prefix over the round's warps + rank inside the warp)."""
    g = off // T
    o = owner_of[off]
    same = [w for w in range(g * T, min(len(owner_of), (g + 1) * T)) if owner_of[w] == o and w < off]
    return header[g] + len(same)


@pytest.mark.parametrize("G,chunk,cap", [(2, 512, 355), (4, 1024, 375), (8, 1024, 221), (3, 700, 300), (8, 1000, 100),
                                         (2, 100, 8), (5, 33, 384)])
def test_sender_packing_equals_receiver_arithmetic(G, chunk, cap):
    rng = np.random.default_rng(G * 1000 + chunk)
    for trial in range(20):
        owner_of = rng.integers(0, G, size=chunk)          # the owner of every walker's partner (uniform partners)
        for me in range(G):
            header, sent, overflow = sender_pack(owner_of, me, chunk, cap)
            for row, off in enumerate(sent):
                assert receiver_row(owner_of, off, header) == row
            for k, off in enumerate(overflow):              # past the slot: the receiver must take the owner-read path
                assert receiver_row(owner_of, off, header) == cap + k >= cap


def task_of(v, G, R, nchunks, lag):
    """Dense ordered sequence number -> ("push", c, slot) | ("update", c, g), exactly the kernel's take():
    region A (chunk index < lag) holds pushes only, region B pushes and the update groups of chunk c - lag, region C
    (chunk index >= nchunks) update groups only -- no sequence number is spent on an empty slot."""
    per_c = G - 1 + R
    len_a, len_b = lag * (G - 1), (nchunks - lag) * per_c
    if v < len_a:
        return ("push",) + divmod(v, G - 1)
    if v < len_a + len_b:
        w = v - len_a
        c, slot = lag + w // per_c, w % per_c
        if slot + 1 < G:
            return ("push", c, slot)
        return ("update", c - lag, slot - (G - 1))
    if v < nchunks * per_c:
        u = (nchunks - lag) * R + (v - len_a - len_b)
        return ("update",) + divmod(u, R)
    return None


@pytest.mark.parametrize("G,R,nchunks,lag", [(2, 2, 40, 7), (8, 4, 33, 33), (4, 4, 17, 1), (3, 1, 9, 3), (1, 4, 12, 5),
                                             (8, 4, 1024, 320), (2, 2, 5, 5)])
def test_task_order_covers_everything_once_and_never_waits_upwards(G, R, nchunks, lag):
    per_c = G - 1 + R
    NT = nchunks * per_c
    assert task_of(NT, G, R, nchunks, lag) is None
    pushes, updates = {}, {}
    for t in range(NT):
        k = task_of(t, G, R, nchunks, lag)
        assert k is not None                                 # one atomic per task: no empty slots
        (pushes if k[0] == "push" else updates)[k[1:]] = t
    assert sorted(pushes) == [(c, s) for c in range(nchunks) for s in range(G - 1)]
    assert sorted(updates) == [(c, g) for c in range(nchunks) for g in range(R)]
    # an update of chunk c waits for the pushes of chunk c on the OTHER ranks, which carry the same sequence numbers
    # there: every one of them precedes the update in the (identical) order -> waits only point downwards; and the
    # updates trail the pushes by `lag` chunks where both run
    for (c, g), tu in updates.items():
        for s in range(G - 1):
            assert pushes[(c, s)] < tu
            if c + lag < nchunks:
                assert pushes[(c + lag, s)] < tu
    # every destination is served exactly once per chunk: dest = (me + 1 + slot) % G for slot in 0..G-2
    for me in range(G):
        assert sorted((me + 1 + s) % G for s in range(G - 1)) == [r for r in range(G) if r != me]


def test_ring_and_flag_indexing_are_injective():
    """Two messages never share a ring slot or a flag: slot (parity, source, chunk), flag (source, chunk)."""
    G, nchunks, cap, d = 4, 6, 10, 2
    slot_bytes = 16 + cap * d * 8
    seen = set()
    for par in range(2):
        for src in range(G):
            for c in range(nchunks):
                lo = ((par * G + src) * nchunks + c) * slot_bytes
                assert all(lo + slot_bytes <= a or b <= lo for a, b in seen)
                seen.add((lo, lo + slot_bytes))
    assert max(b for _, b in seen) == 2 * G * nchunks * slot_bytes          # what the host allocates for the ring
    assert len({src * nchunks + c for src in range(G) for c in range(nchunks)}) == G * nchunks


def test_two_ring_parities_suffice_under_the_stated_ordering():
    """Model check of the claim in kmc_push.cuh: rank q writes ring parity h & 1 for half-step h only after every rank
    finished update h - 1; a rank finishes update h only after it consumed every message of half-step h.  So when q
    writes parity p again (half-step h + 2), every consumer has finished update h + 1 > h: the slot's previous content
    (half-step h) is dead.  Exhaustive over a small event trace."""
    G, H = 3, 6
    done_update = [-1] * G                      # last half-step whose updates a rank has finished
    for h in range(H):
        for q in range(G):                      # pushes of half-step h: allowed once everybody finished update h - 1
            assert all(du >= h - 1 for du in done_update)
            for r in range(G):
                if r != q:                      # overwrites parity h & 1 at r: last written for half-step h - 2
                    assert done_update[r] >= h - 2
        for r in range(G):
            done_update[r] = h
