// kmc_push.cuh -- the sharded ensemble (one ensemble over G GPUs) as ONE persistent kernel
// per GPU with OWNER-COMPUTES PUSHES of the partner rows over NVLink.
//
// Replaces, across GPUs, the reference's shared-memory read of the passive half inside the threaded sweep
// (src/samplers.jl:246-273, partner read at :255): rank r updates positions [r*S, (r+1)*S) of each half and holds
// ONLY those rows (x[2][S][D]); the partner row of an active walker lives on rank (partner position / S).
//
// The draws are counter-based (Philox keyed by the global walker index, kmc_device.cuh), so the OWNER of a passive
// shard can enumerate, without any index round trip, which of its rows every other rank's active walkers will ask for
// (SURVEY.md section 7.2(9)(i)).  Per half-step every rank runs two kinds of tasks, handed out to its CTAs from one
// atomic counter in a fixed order:
//
//   push(c, dest)   enumerate the partner draws of chunk c (`chunk` consecutive active walkers) of rank `dest`; the
//                   hits in my passive shard are gathered (one bulk copy per row), PACKED in walker order in shared
//                   memory and sent as ONE bulk store (cp.async.bulk shared -> peer global, up to cap*8D bytes) into
//                   dest's receive ring slot (half-step parity, source = me, chunk c); when the store has completed,
//                   the "chunk ready" flag (source, c) in dest's memory is set to h+1 (st.release.sys).
//   update(c)       the stretch-move step of my chunk c: the same draws give every walker's owner and its rank among
//                   the chunk's walkers with that owner = its row in the packed message; waits for the G-1 chunk
//                   flags, then runs the bulk kernel's group loop (kmc_kernels.cuh, emcee_bulk_kernel) with the
//                   partner row gathered from the local receive ring (remote owner), the local passive shard (own
//                   rows) or, for the rare row past the ring slot's capacity, straight from the owner's memory.
//
// update(c) is handed out `lag` chunks after push(c, *), so transfers are in flight while earlier chunks are updated;
// there is no cross-GPU barrier at all -- a consumer starts as soon as ITS chunk's rows have landed -- and one local
// grid barrier per half-step (random rows of the whole shard are read by the next half-step's pushes).
//
// Why there is no deadlock: give push(c, *) level c and update(c) level c + lag + 1/2.  Tasks are taken in level order
// by co-resident CTAs (cooperative launch), a task only ever waits for strictly lower levels of the same half-step on
// other GPUs, and pushes wait for nothing.  Why two ring parities suffice: rank q can only push for half-step h+2 after
// it finished update h+1, which needs every rank's pushes of h+1, which a rank sends only after its update h.
//
// Exactness: every walker-step is the same arithmetic on the same draws as the single-GPU kernels, so a sharded run is
// bit-identical to the single-GPU run of the same ensemble (tests/test_gpu_push.py).
#pragma once
#include "kmc_kernels.cuh"

namespace kmc {

constexpr int kPushThreads = 256;
constexpr int kPushMaxRounds = 4;  // chunk <= 4 * 256 walkers
constexpr int kPushMaxRanks = 8;
constexpr int kPushSlots = 32;     // (round, warp) slots of a chunk: 4 rounds x 8 warps
static_assert(kPushSlots * kPushMaxRanks == kPushThreads, "one thread zeroes one counter");
static_assert(kPushMaxRounds * (kPushThreads / 32) == kPushSlots, "slots = rounds x warps = one warp's lanes");

struct PushParams {
    double *recv;               // local receive ring [2 parities][G sources][nchunks][cap][D]
    unsigned long long *flags;  // local chunk flags [G sources][nchunks]: h+1 once the rows for half-step h have landed
    double *peer_recv[kPushMaxRanks];
    unsigned long long *peer_flags[kPushMaxRanks];
    const double *peer_x[kPushMaxRanks];  // x[2][S][D] of every rank (capacity-overflow fallback reads)
    unsigned long long *task_ctr;          // task counter (zeroed by the host before every launch)
    unsigned S;        // shard size: positions of each half per rank
    unsigned G, rank;  // ranks, my rank
    unsigned chunk;    // walkers per chunk (<= 1024)
    unsigned rounds;   // ceil(chunk / 256)
    unsigned nchunks;  // ceil(S / chunk)
    unsigned cap;      // rows per ring slot (<= 256)
    unsigned lag;      // update(c) follows push(c + lag, *)
};

// The partner draw alone (src/samplers.jl:250): the owner side of a push needs nothing else of the walker-step.
__device__ __forceinline__ unsigned partner_pos(const RunParams &p, long long h, unsigned i) {
    const unsigned long long t = (unsigned long long)(h >> 1);
    const unsigned batch = (unsigned)(h & 1);
    const unsigned walker = p.id_base[batch] + i;
    const Philox4 r = philox4x32_10(walker, (unsigned)t, (unsigned)(t >> 32), batch, p.keys);
    unsigned hi, lo;
    mul_wide(r.r0, p.nhalf, hi, lo);
    if (lo < p.lemire_t) hi = lemire_retry(p.keys, walker, (unsigned)t, (unsigned)(t >> 32), batch, p.nhalf, p.lemire_t);
    return hi;  // position inside the passive half
}

__device__ __forceinline__ void flag_publish(unsigned long long *flag, unsigned long long v) {
    asm volatile("fence.proxy.async;" ::: "memory");  // the bulk store's writes (async proxy) before the flag (generic proxy)
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(v) : "memory");
}

__device__ __forceinline__ void flag_wait(const unsigned long long *flag, unsigned long long v) {
    unsigned long long cur;
    long long t0 = 0;
    unsigned spins = 0;
    do {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(cur) : "l"(flag) : "memory");
        if (cur < v && (++spins & 0x3FFu) == 0) {  // watchdog (~10 s): a missing peer must not hang the GPU
            if (t0 == 0) t0 = clock64();
            else if (clock64() - t0 > 20000000000LL) __trap();
        }
    } while (cur < v);
}

template <template <int> class Dn, int D>
__global__ void __launch_bounds__(kPushThreads, 3) emcee_push_kernel(const RunParams p, const PushParams q,
                                                                      const Dn<D> dn) {
    static_assert(D % 2 == 0, "rows must be multiples of 16 bytes");
    extern __shared__ __align__(128) unsigned char push_smem[];
    constexpr unsigned T = kPushThreads, ROWB = D * 8;
    double *buf = reinterpret_cast<double *>(push_smem);  // [2][T][D]: own rows of an update group | packed rows of a push
    double *xpart = buf + 2 * T * D;                      // [T][D] partner rows of an update group
    unsigned long long *ldbar = reinterpret_cast<unsigned long long *>(xpart + T * D);
    unsigned long long *next_slot = ldbar + 1;            // broadcast of the next task id
    unsigned *cnt = reinterpret_cast<unsigned *>(ldbar + 2);  // [kPushSlots][kPushMaxRanks] hits per (round, warp) and owner
    unsigned *pre = cnt + kPushSlots * kPushMaxRanks;         // exclusive prefix of cnt over the slots, per owner
    unsigned *tot = pre + kPushSlots * kPushMaxRanks;         // [kPushMaxRanks] totals
    unsigned short *rowidx = reinterpret_cast<unsigned short *>(tot + kPushMaxRanks);  // [chunk] row in the packed message

    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned S = q.S, me = q.rank, G = q.G;
    const unsigned NT = (q.nchunks + q.lag) * G;  // tasks per half-step
    if (tid == 0) kbar_init(ldbar, 1);
    __syncthreads();
    unsigned ldphase = 0, unit = 0;
    // thread 0: chunk flags of the last two store units, published once their bulk store has completed
    // (scalars selected by parity, not arrays: dynamic indexing would put them in local memory)
    unsigned long long *pf0 = nullptr, *pf1 = nullptr;
    unsigned long long pv0 = 0, pv1 = 0;
    auto publish_pending = [&](bool first, bool second) {
        if (first && pf0) {
            flag_publish(pf0, pv0);
            pf0 = nullptr;
        }
        if (second && pf1) {
            flag_publish(pf1, pv1);
            pf1 = nullptr;
        }
    };

    // A "unit" is a push task or an update group: it owns buf[unit & 1] and commits exactly one bulk store group.
    // Its buffer was last used two units ago: all but the latest store group must be complete (which also lets the
    // flag of the unit two back go out), then everyone may overwrite the buffer (and xpart).
    auto unit_begin = [&]() -> double * {
        if (tid == 0) {
            asm volatile("cp.async.bulk.wait_group 1;" ::: "memory");
            publish_pending((unit & 1) == 0, (unit & 1) == 1);
        }
        __syncthreads();
        return buf + (size_t)(unit & 1) * T * D;
    };

    auto grab = [&]() -> unsigned long long {  // thread 0 only
        return atomicAdd(q.task_ctr, 1ULL);
    };

#ifdef KMC_PUSH_PROF  // thread 0's cycles per phase (experiment builds only)
    long long pt[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, pc0 = clock64();
    unsigned npush = 0, nupd = 0;
#define PUSH_TICK(i) do { const long long c_ = clock64(); pt[i] += c_ - pc0; pc0 = c_; } while (0)
#else
#define PUSH_TICK(i) do { } while (0)
#endif
    long long n = p.n0, phase = p.phase0, sidx = p.sidx0;
    unsigned long long target = p.bar_base;
    unsigned long long tbeg = 0;
    unsigned long long next = 0;
    if (tid == 0) *next_slot = grab();
    __syncthreads();
    next = *next_slot;

    for (long long h = p.h0; h < p.h1; ++h) {
        const unsigned batch = (unsigned)(h & 1);
        const bool store = (n > 0) && (phase == 0);      // :268
        const bool reset = (batch == 1) && (n == 0);     // :285-288
        const size_t act = batch ? (size_t)S : 0;         // first local row of the active half
        const size_t pas = batch ? 0 : (size_t)S;         // first local row of the passive half
        const unsigned par = (unsigned)(h & 1);           // ring parity
        const unsigned long long ready = (unsigned long long)h + 1;
        const unsigned long long tend = tbeg + NT;

        while (next < tend) {
            const unsigned t = (unsigned)(next - tbeg);
            unsigned long long nxt = 0;
            if (tid == 0) nxt = grab();  // the atomic's round trip hides behind this task
            const unsigned c = t / G, slot = t - c * G;
            if (slot + 1 < G) {
                // ------------------------------------------------------------ push(c, dest)
                if (c < q.nchunks) {
                    const unsigned dest = (me + 1 + slot) % G;
                    const unsigned i0 = dest * S + c * q.chunk;  // first active walker (position in its half) of the chunk
                    const unsigned lim = min(q.chunk, S - c * q.chunk);
                    PUSH_TICK(9);
                    double *pk = unit_begin();
                    PUSH_TICK(0);
                    unsigned lrow[kPushMaxRounds], rk[kPushMaxRounds];
                    if (tid < kPushSlots) cnt[tid] = 0u;
                    __syncthreads();
#pragma unroll
                    for (int g = 0; g < kPushMaxRounds; ++g) {
                        lrow[g] = 0xFFFFFFFFu;
                        rk[g] = 0;
                        if (g < (int)q.rounds) {
                            const unsigned off = g * T + tid;
                            bool hit = false;
                            if (off < lim) {
                                const unsigned pl = partner_pos(p, h, i0 + off) - me * S;  // < S iff the partner is mine
                                hit = pl < S;
                                if (hit) lrow[g] = pl;
                            }
                            const unsigned b = __ballot_sync(0xffffffffu, hit);
                            rk[g] = __popc(b & ((1u << lane) - 1u));
                            if (lane == 0) cnt[g * 8 + warp] = __popc(b);
                        }
                    }
                    __syncthreads();
                    if (warp == 0) {
                        const unsigned v = cnt[lane];
                        unsigned incl = v;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const unsigned u = __shfl_up_sync(0xffffffffu, incl, o);
                            if ((int)lane >= o) incl += u;
                        }
                        pre[lane] = incl - v;
                        if (lane == 31) tot[0] = incl;
                    }
                    __syncthreads();
                    const unsigned nsend = min(tot[0], q.cap);
                    PUSH_TICK(1);
                    if (tid == 0) kbar_expect_tx(ldbar, nsend * ROWB);
#pragma unroll
                    for (int g = 0; g < kPushMaxRounds; ++g) {
                        if (lrow[g] != 0xFFFFFFFFu) {
                            const unsigned pi = pre[g * 8 + warp] + rk[g];
                            if (pi < q.cap) bulk_row_g2s(pk + (size_t)pi * D, p.x + (pas + lrow[g]) * D, ROWB, ldbar);
                        }
                    }
                    kbar_wait(ldbar, ldphase);
                    ldphase ^= 1;
                    PUSH_TICK(2);
                    if (tid == 0) {
                        if (nsend) {
                            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                            double *dst = q.peer_recv[dest] + (((size_t)par * G + me) * q.nchunks + c) * q.cap * D;
                            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
                                         "r"((unsigned)__cvta_generic_to_shared(pk)), "r"(nsend * ROWB)
                                         : "memory");
                        }
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        unsigned long long *fl = q.peer_flags[dest] + (size_t)me * q.nchunks + c;
                        if (unit & 1) {
                            pf1 = fl;
                            pv1 = ready;
                        } else {
                            pf0 = fl;
                            pv0 = ready;
                        }
                    }
                    ++unit;
                    PUSH_TICK(3);
#ifdef KMC_PUSH_PROF
                    ++npush;
#endif
                }
            } else if (c >= q.lag) {
                // ------------------------------------------------------------ update(c - lag)
                const unsigned cu = c - q.lag;
                const unsigned l0 = cu * q.chunk;            // first local position of the chunk
                const unsigned lim = min(q.chunk, S - l0);
                // pass 1: owner and packed-message row of every walker's partner
                PUSH_TICK(9);
                unsigned char own8[kPushMaxRounds], rk8[kPushMaxRounds];
                // No flag of mine may stay unpublished while I wait for somebody else's (two CTAs on two GPUs could
                // otherwise wait for each other's deferred flags): complete my stores and publish first.
                if (tid == 0 && (pf0 || pf1)) {
                    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                    publish_pending(true, true);
                }
                PUSH_TICK(4);
                cnt[tid] = 0u;  // kPushSlots * kPushMaxRanks == T (the previous task ended with a CTA barrier)
                __syncthreads();
#pragma unroll
                for (int g = 0; g < kPushMaxRounds; ++g) {
                    own8[g] = 0xFF;
                    rk8[g] = 0;
                    if (g < (int)q.rounds) {
                        const unsigned off = g * T + tid;
                        unsigned owner = 0xFFu;
                        if (off < lim) owner = partner_pos(p, h, me * S + l0 + off) / S;
                        own8[g] = (unsigned char)owner;
                        for (unsigned o = 0; o < G; ++o) {
                            const unsigned b = __ballot_sync(0xffffffffu, owner == o);
                            if (owner == o) rk8[g] = (unsigned char)__popc(b & ((1u << lane) - 1u));
                            if (lane == 0) cnt[(g * 8 + warp) * kPushMaxRanks + o] = __popc(b);
                        }
                    }
                }
                __syncthreads();
                if (warp == 0) {
                    for (unsigned o = 0; o < G; ++o) {
                        const unsigned v = cnt[lane * kPushMaxRanks + o];
                        unsigned incl = v;
#pragma unroll
                        for (int k = 1; k < 32; k <<= 1) {
                            const unsigned u = __shfl_up_sync(0xffffffffu, incl, k);
                            if ((int)lane >= k) incl += u;
                        }
                        pre[lane * kPushMaxRanks + o] = incl - v;
                    }
                    // every source's rows of this chunk have landed in my ring (their flags were set after their stores)
                    PUSH_TICK(5);
                    if (lane < G && lane != me) flag_wait(q.flags + (size_t)lane * q.nchunks + cu, ready);
                    PUSH_TICK(6);
                }
                __syncthreads();
#pragma unroll
                for (int g = 0; g < kPushMaxRounds; ++g)
                    if (own8[g] != 0xFF)
                        rowidx[g * T + tid] = (unsigned short)(pre[(g * 8 + warp) * kPushMaxRanks + own8[g]] + rk8[g]);
                asm volatile("fence.proxy.async;" ::: "memory");  // acquired peer writes -> this thread's bulk gathers

                // pass 2: the walker-steps, in groups of T (emcee_bulk_kernel's group loop)
                const double *ring = q.recv + (size_t)par * G * q.nchunks * q.cap * D;
                for (unsigned g = 0; g * T < lim; ++g) {
                    const unsigned rows = min(T, lim - g * T);
                    const unsigned l = l0 + g * T + tid;  // local position
                    const bool live = tid < rows;
                    double *ownb = unit_begin();
                    if (tid == 0) {
                        kbar_expect_tx(ldbar, rows * ROWB * 2);
                        bulk_row_g2s(ownb, p.x + (act + l0 + (size_t)g * T) * D, rows * ROWB, ldbar);
                    }
                    DrawRec dr;
                    dr.j = 0;
                    dr.z = 0.0;
                    dr.q = 0.f;
                    if (live) {
                        unsigned j;
                        double z, u;
                        step_draws<false>(p, h, me * S + l, j, z, u);  // :250, :252, (:260 uniform)
                        dr.z = z;
                        dr.q = filter_q<false>(p, z, u);
                        const unsigned pl = j >= p.nhalf ? j - p.nhalf : j;  // position inside the passive half
                        const unsigned owner = pl / S, prow = pl - owner * S;
                        const double *src;
                        if (owner == me) {
                            src = p.x + (pas + prow) * D;
                        } else {
                            const unsigned pi = rowidx[g * T + tid];
                            if (pi < q.cap) src = ring + (((size_t)owner * q.nchunks + cu) * q.cap + pi) * D;
                            else src = q.peer_x[owner] + (pas + prow) * D;  // past the slot's capacity: read the owner
                        }
                        bulk_row_g2s(xpart + (size_t)tid * D, src, ROWB, ldbar);
                    }
                    const size_t k = act + l;
                    const double lpk = live ? p.lp[k] : 0.0;
                    kbar_wait(ldbar, ldphase);
                    ldphase ^= 1;
                    if (live) {
                        double xk[D], xj[D], y[D];
#pragma unroll
                        for (int cc = 0; cc < D; cc += 2) {
                            const double2 a2 = *reinterpret_cast<const double2 *>(ownb + (size_t)tid * D + cc);
                            const double2 b2 = *reinterpret_cast<const double2 *>(xpart + (size_t)tid * D + cc);
                            xk[cc] = a2.x;
                            xk[cc + 1] = a2.y;
                            xj[cc] = b2.x;
                            xj[cc + 1] = b2.y;
                        }
                        const double z = dr.z;
#pragma unroll
                        for (int cc = 0; cc < D; ++cc) y[cc] = dadd(xj[cc], dmul(z, dsub(xk[cc], xj[cc])));  // :255
                        const double p1 = dn.logpdf(y);                                                    // :257
                        const double tt = (p1 - lpk) + (double)dr.q * 0.6931471805599453;                 // :260
                        bool acc;
                        if (tt > (double)p.margin) acc = true;
                        else if (tt < -(double)p.margin) acc = false;
                        else acc = accept_slow<false, false>(p, h, me * S + l, z, p1, lpk);
                        if (acc) {  // :261-265
#pragma unroll
                            for (int cc = 0; cc < D; cc += 2)
                                *reinterpret_cast<double2 *>(ownb + (size_t)tid * D + cc) = make_double2(y[cc], y[cc + 1]);
                            p.lp[k] = p1;
                            if (!reset) p.nacc[k] += 1u;
                        }
                        if (reset) {  // :285-288 burn-in counters are discarded (both halves of this position)
                            p.nacc[l] = 0u;
                            p.nacc[(size_t)S + l] = 0u;
                        }
                        if (store) chain_store<D>(p, chain_row(p, sidx, batch, me * S + l), acc, y, xk, p1, lpk);
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // smem writes -> bulk store
                    __syncthreads();
                    if (tid == 0) bulk_s2g(p.x + (act + l0 + (size_t)g * T) * D, ownb, rows * ROWB);
                    ++unit;
                }
                PUSH_TICK(7);
#ifdef KMC_PUSH_PROF
                ++nupd;
#endif
            }
            __syncthreads();  // everyone has read `next` of this iteration before it is overwritten
            if (tid == 0) *next_slot = nxt;
            __syncthreads();
            next = *next_slot;
            PUSH_TICK(8);
        }

        if (batch == 1) {
            if (store) ++sidx;
            ++n;
            if (++phase == p.nthin) phase = 0;
        }
        if (tid == 0) {  // all of this CTA's stores are complete: the last flags go out, own rows are final
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            publish_pending(true, true);
            asm volatile("fence.proxy.async;" ::: "memory");
        }
        tbeg = tend;
        if (h + 1 < p.h1) {  // the reference's join between the two sweeps (:248/:273), local to this GPU
            target += gridDim.x;
            __syncthreads();
            if (gridDim.x > 1) {
                if (tid == 0) {
                    barrier_arrive(p.barrier);
                    barrier_wait(p.barrier, target);
                }
                __syncthreads();
            }
        }
        PUSH_TICK(9);
    }
#ifdef KMC_PUSH_PROF
    if (tid == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x / 2))
        printf("push rank %u cta %d: pushes %u updates %u | cycles: unit_begin %lld enum+scan %lld gather %lld send %lld | "
               "flush %lld pass1 %lld flagwait %lld groups %lld | handoff %lld barrier+other %lld\n",
               me, (int)blockIdx.x, npush, nupd, pt[0], pt[1], pt[2], pt[3], pt[4], pt[5], pt[6], pt[7], pt[8], pt[9]);
#endif
}

}  // namespace kmc
