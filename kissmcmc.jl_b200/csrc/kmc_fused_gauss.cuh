// kmc_fused_gauss.cuh -- K2F: the whole stretch-move half-step of the dense Gaussian plugin
// (BASELINE.json configs[2], 16 < d <= 128) as ONE persistent tcgen05 kernel.
//
// One CTA per SM, 512 threads, cooperative launch over a whole range of half-steps (grid barrier
// between half-steps like emcee_run_kernel).  The matrix pieces A (3 x 32 KB bf16, SWIZZLE_128B)
// are loaded once by TMA and stay in shared memory.  Per 128-walker tile:
//
//   P1  all 16 warps, 16 lanes per walker (two walkers in flight per half-warp), 8 columns per lane (the draws,
//       src/samplers.jl:250,:252,:260, were made one thread per walker while the previous tile's GEMM ran):
//       gather x_j (one contiguous 8d-byte row) and x_k, proposal y = xj + z(xk - xj) (:255) in FP64,
//       centre, split into three bf16 pieces and store each 16-byte chunk DIRECTLY into the swizzled
//       K-major UMMA tile in shared memory (chunk index XOR row%8 inside each 8 x 128 B atom) --
//       the proposal never goes to global memory.   fence.proxy.async + bar.sync
//   P2  one thread issues the 48 tcgen05.mma (6 piece pairs x 8 k-steps, M=128 N=128 K=16, FP32 in
//       TMEM) and commits to an mbarrier; warps 0-3 (thread = walker row) read the accumulator with
//       tcgen05.ld, reduce |y|^2 over the first d columns, form logp and run the exact FP64 accept
//       test (:260), update logp / accept counter / thinned logp chain
//   P3  all warps: accepted walkers (and all, when the iteration is stored) recompute y with the
//       same three IEEE operations and write x (:261) / the chain (:268-272)
//
// Traffic per half-step ~ read x_k, x_j once + rewrite accepted rows: close to the algorithmic
// 24d+24 bytes per walker-step, instead of ~3x that for the three-kernel pipeline.
#pragma once
#include "kmc_batched.cuh"
#include "kmc_tc.cuh"

namespace kmc {
namespace tc {

constexpr int kFusedThreads = 512;

struct __align__(1024) FusedSmem {
    unsigned char a[PIECES][GPIECE_BYTES];  // matrix pieces (B operand), resident
    unsigned char c[PIECES][GPIECE_BYTES];  // centred proposal pieces (A operand) of the current tile
    unsigned long long afull, mma_done;
    unsigned tmem_base;
    double z[2][BM], u[2][BM];   // per-row draws, double-buffered by tile parity (P3 of tile t overlaps P1 of tile t+1)
    unsigned j[2][BM];
    unsigned char acc[2][BM];
};

struct FusedParams {
    const double *mu;  // [d] (device)
    double lognorm;
    int d;
};

// byte offset of the 16-byte chunk (row r, chunk index ck in 0..15 = 8 bf16 columns each) inside a
// piece stored as [2 k-halves][128 rows][128 B] with the 128-byte swizzle
__device__ __forceinline__ unsigned sw128_chunk_offset(unsigned r, unsigned ck) {
    const unsigned kh = ck >> 3, c8 = ck & 7;
    return kh * (GPIECE_BYTES / 2) + (r >> 3) * 1024 + (r & 7) * 128 + ((c8 ^ (r & 7)) << 4);
}

template <bool REPLAY>
__global__ void __launch_bounds__(kFusedThreads, 1)
gaussian_fused_kernel(const __grid_constant__ CUtensorMap mapA, const RunParams p, const FusedParams fp) {
    extern __shared__ unsigned char smem_raw[];
    FusedSmem &sm = *reinterpret_cast<FusedSmem *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int d = fp.d;

    if (tid == 0) {
        mbar_init(&sm.afull, 1);
        mbar_init(&sm.mma_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&sm.tmem_base))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem = sm.tmem_base;
    if (tid == 0) {  // the matrix: once per CTA
        mbar_expect_tx(&sm.afull, PIECES * GPIECE_BYTES);
        for (int pc = 0; pc < PIECES; ++pc)
            for (int kh = 0; kh < 2; ++kh)
                tma_load_2d(sm.a[pc] + kh * (GPIECE_BYTES / 2), &mapA, kh * 64, pc * GN, &sm.afull);
    }
    mbar_wait(&sm.afull, 0);

    const unsigned W = p.shard_end - p.shard_begin;
    const unsigned ntiles = (W + BM - 1) / BM;
    const unsigned half16 = lane >> 4;   // which of the warp's two walkers
    const unsigned ck = lane & 15;       // 16-byte chunk = columns 8*ck .. 8*ck+7
    unsigned mma_phase = 0, tpar = 0;  // tpar: parity of the tile counter of this CTA
    long long n = p.n0, phase = p.phase0, sidx = p.sidx0;
    unsigned long long target = p.bar_base;

    for (long long h = p.h0; h < p.h1; ++h) {
        const unsigned batch = (unsigned)(h & 1);
        const bool store = (n > 0) && (phase == 0);  // :268
        const size_t a0 = batch ? (size_t)p.nhalf : 0;

        // draws (src/samplers.jl:250,:252,:260) of one tile: one thread per walker row, into parity `par`
        auto tile_draws = [&](unsigned tl, unsigned par, unsigned r) {
            const unsigned w = tl * BM + r;
            if (w < W) {
                unsigned j;
                double z, u;
                step_draws<REPLAY>(p, h, p.shard_begin + w, j, z, u);
                sm.z[par][r] = z;
                sm.u[par][r] = u;
                sm.j[par][r] = j;
            }
        };
        if (blockIdx.x < ntiles && tid < BM) tile_draws(blockIdx.x, tpar, tid);
        __syncthreads();
        for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const unsigned w0 = tile * BM;
            // ------------------------------------------------ P1: proposals -> swizzled bf16 pieces
            // lane handles column pairs cp = ck + 16 e (columns 2cp, 2cp+1), e = 0..3: every load instruction of
            // a half-warp reads 256 contiguous bytes of the row
            // Two rows per half-warp are in flight at once (rows r and r + 64): all loads first, then the math.
            for (unsigned rb = warp * 2 + half16; rb < BM / 2; rb += 2 * (kFusedThreads / 32)) {
                double2 xa[2][4], xb[2][4];
                double zz[2];
                bool live[2];
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    const unsigned r = rb + t * (BM / 2), w = w0 + r;
                    live[t] = w < W;
                    zz[t] = 0.0;
                    if (live[t]) {
                        const unsigned i = p.shard_begin + w;
                        const unsigned j = sm.j[tpar][r];
                        zz[t] = sm.z[tpar][r];
                        const double *xk = p.x + (a0 + i) * d, *xj = p.x + (size_t)j * d;
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int c = 2 * (ck + 16 * e);
                            xa[t][e] = make_double2(0.0, 0.0);
                            xb[t][e] = make_double2(0.0, 0.0);
                            if ((d & 1) == 0) {
                                if (c < d) {
                                    xa[t][e] = *reinterpret_cast<const double2 *>(xk + c);
                                    xb[t][e] = __ldcg(reinterpret_cast<const double2 *>(xj + c));
                                }
                            } else {
                                if (c < d) {
                                    xa[t][e].x = xk[c];
                                    xb[t][e].x = __ldcg(xj + c);
                                }
                                if (c + 1 < d) {
                                    xa[t][e].y = xk[c + 1];
                                    xb[t][e].y = __ldcg(xj + c + 1);
                                }
                            }
                        }
                    }
                }
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    const unsigned r = rb + t * (BM / 2);
                    double v[8];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int c = 2 * (ck + 16 * e);
                        v[2 * e] = (live[t] && c < d) ? dadd(xb[t][e].x, dmul(zz[t], dsub(xa[t][e].x, xb[t][e].x))) - fp.mu[c] : 0.0;  // :255, centred
                        v[2 * e + 1] = (live[t] && c + 1 < d) ? dadd(xb[t][e].y, dmul(zz[t], dsub(xa[t][e].y, xb[t][e].y))) - fp.mu[c + 1] : 0.0;
                    }
#pragma unroll
                    for (int pc = 0; pc < PIECES; ++pc) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const __nv_bfloat16 h0 = __double2bfloat16(v[2 * e]), h1 = __double2bfloat16(v[2 * e + 1]);
                            v[2 * e] -= (double)__bfloat162float(h0);
                            v[2 * e + 1] -= (double)__bfloat162float(h1);
                            const unsigned cp = ck + 16 * e;  // column pair -> 4 bytes inside chunk cp/4
                            *reinterpret_cast<unsigned *>(sm.c[pc] + sw128_chunk_offset(r, cp >> 2) + ((cp & 3) << 2)) =
                                (unsigned)__bfloat16_as_ushort(h0) | ((unsigned)__bfloat16_as_ushort(h1) << 16);
                        }
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic smem writes -> tensor-core proxy
            __syncthreads();
            // ------------------------------------------------ P2: tcgen05 GEMM, epilogue, accept
            if (tid == 0) {
                tc_fence_after();
                constexpr unsigned idesc = idesc_bf16_f32(BM, GN);
                const int pc_c[6] = {2, 0, 1, 1, 0, 0};
                const int pc_a[6] = {0, 2, 1, 0, 1, 0};
#pragma unroll
                for (int pr = 0; pr < 6; ++pr) {
                    const unsigned cb = smem_u32(sm.c[pc_c[pr]]), ab = smem_u32(sm.a[pc_a[pr]]);
#pragma unroll
                    for (int k = 0; k < GK / 16; ++k) {
                        const unsigned off = (k >> 2) * (GPIECE_BYTES / 2) + (k & 3) * 32;
                        tc_mma(tmem, smem_desc_sw128(cb + off), smem_desc_sw128(ab + off), idesc, (pr | k) ? 1u : 0u);
                    }
                }
                tc_commit(&sm.mma_done);
            }
            if (warp < 4) {
                mbar_wait(&sm.mma_done, mma_phase);
                __syncwarp();  // lane 0 came here from the MMA issue: converge before the .aligned TMEM loads
                tc_fence_after();
                const int r = warp * 32 + lane;
                double ss = 0.0;
#pragma unroll 1
                for (int cb = 0; cb < GN; cb += 32) {
                    if (cb >= d) break;
                    unsigned v[32];
                    tmem_ld32(tmem + ((unsigned)(warp * 32) << 16) + cb, v);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    float part = 0.0f;
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        const float y = __uint_as_float(v[e]);
                        part = fmaf(y, y, part);
                    }
                    ss += (double)part;
                }
                tc_fence_before();
                const unsigned w = w0 + r;
                if (w < W) {
                    const unsigned i = p.shard_begin + w;
                    const size_t k = a0 + i;
                    const double p1 = fp.lognorm - 0.5 * ss, p0 = p.lp[k];
                    const bool acc = accept_exact<false>(p.nm1, sm.z[tpar][r], p1, p0, sm.u[tpar][r]);  // :260
                    sm.acc[tpar][r] = acc ? 1 : 0;
                    if (acc) {
                        p.lp[k] = p1;
                        p.nacc[k] += 1u;
                    }
                    if (batch == 1 && n == 0) {  // :285-288
                        p.nacc[i] = 0u;
                        p.nacc[(size_t)p.nhalf + i] = 0u;
                    }
                    if (store) __stcs(p.chain_lp + chain_row(p, sidx, batch, i), acc ? p1 : p0);
                }
            }
            else if (warp < 8) {  // idle during the GEMM: the NEXT tile's draws, into the other parity
                if (tile + gridDim.x < ntiles) tile_draws(tile + gridDim.x, tpar ^ 1, (warp - 4) * 32 + lane);
            }
            mma_phase ^= 1;
            __syncthreads();
            // ------------------------------------------------ P3: accepted rows (and the chain)
            // No barrier after P3: the next tile's P1 only writes sm.c (free since the MMA completed) and the
            // OTHER parity of the per-row arrays; the bar.sync after that P1 orders everything else.
            for (unsigned r = warp * 2 + half16; r < BM; r += 2 * (kFusedThreads / 32)) {
                const unsigned w = w0 + r;
                if (w >= W) continue;
                const bool accr = sm.acc[tpar][r] != 0;
                if (!accr && !store) continue;
                const unsigned i = p.shard_begin + w;
                double *xk = p.x + (a0 + i) * d;
                const double *xj = p.x + (size_t)sm.j[tpar][r] * d;
                const double z = sm.z[tpar][r];
                const size_t o = store ? chain_row(p, sidx, batch, i) : 0;
                double2 xa[4], xb[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {  // all loads of the row first
                    const int c = 2 * (ck + 16 * e);
                    xa[e] = make_double2(0.0, 0.0);
                    xb[e] = make_double2(0.0, 0.0);
                    if ((d & 1) == 0) {
                        if (c < d) {
                            xa[e] = *reinterpret_cast<const double2 *>(xk + c);
                            if (accr) xb[e] = __ldcg(reinterpret_cast<const double2 *>(xj + c));
                        }
                    } else {
                        if (c < d) {
                            xa[e].x = xk[c];
                            if (accr) xb[e].x = __ldcg(xj + c);
                        }
                        if (c + 1 < d) {
                            xa[e].y = xk[c + 1];
                            if (accr) xb[e].y = __ldcg(xj + c + 1);
                        }
                    }
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int c = 2 * (ck + 16 * e);
                    const double v0 = accr ? dadd(xb[e].x, dmul(z, dsub(xa[e].x, xb[e].x))) : xa[e].x;  // :255, same bits
                    const double v1 = accr ? dadd(xb[e].y, dmul(z, dsub(xa[e].y, xb[e].y))) : xa[e].y;
                    if (c < d) {
                        if (accr) xk[c] = v0;                            // :261
                        if (store) __stcs(p.chain_x + o * d + c, v0);   // :268-272
                    }
                    if (c + 1 < d) {
                        if (accr) xk[c + 1] = v1;
                        if (store) __stcs(p.chain_x + o * d + c + 1, v1);
                    }
                }
            }
            tpar ^= 1;
        }
        if (batch == 1) {
            if (store) ++sidx;
            ++n;
            if (++phase == p.nthin) phase = 0;
        }
        if (h + 1 < p.h1) {  // the reference's join between the two half-ensemble sweeps (:248/:273)
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
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}

}  // namespace tc
}  // namespace kmc
