// Weighted covariance  V_k[bin] = (1/T) sum_t phi_k(t) x(t) x(t)^H  for all K sources in ONE pass over X.
// (reference: overiva.py:179 -- K separate zgemm calls, each streaming X and an |X|-sized temporary;
//  overiva.py:87 for the unweighted input covariance.)
//
// Mapping (DESIGN.md "covariance kernel"): lane <-> frequency bin.
//  * a "team" of P warps owns one group of 32 bins at a time.  The group's frames arrive in shared memory
//    in chunks of TC frames ([t][c][32 lanes] complex, one contiguous block = one 1-D bulk-TMA transaction)
//    through a ring of S stages; the team's leader lane issues the copies S-1 chunks ahead
//    (cp.async.bulk + mbarrier complete_tx), consumers hand stages back through an "empty" mbarrier.
//    The phi weights of the chunk ride in the same stage.
//  * each lane reads ITS bin's M channel values of a frame (one conflict-free LDS.128 per channel), forms
//    the Hermitian products p_ij = x_i conj(x_j) once for all sources and accumulates phi_k * p_ij in
//    registers.  The accumulators of a bin never leave their lane: there is no cross-lane reduction.
//    When the lower triangle does not fit one lane's registers it is split over the P warps of the team
//    (entry e belongs to part e % P; a warp's part is compile-time, so register indexing is static).
//  * at the end of the group the lane writes its entries, scaled by 1/T, to the grouped lower-triangle
//    layout Vg[gi][k][e][lane] (512-byte coalesced stores).  Few-groups / long-mixture inputs split the
//    frames of a group over several teams; each split writes its partial sum to its own slot of a scratch buffer
//    and a second tiny kernel adds the slots in a fixed order (bit-reproducible), or, without scratch, the partial
//    sums are added atomically into a zeroed Vg.
#pragma once
#include <type_traits>

#include "common.cuh"

namespace oiva {

template <int N, int I = 0, typename F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<N, I + 1>(f);
    }
}

__host__ __device__ constexpr int ent_row(int e) {
    int i = 0;
    while ((i + 1) * (i + 2) / 2 <= e) ++i;
    return i;
}
__host__ __device__ constexpr int ent_col(int e) { return e - ent_row(e) * (ent_row(e) + 1) / 2; }

#ifndef OIVA_COV_REG_BUDGET
#define OIVA_COV_REG_BUDGET 100  // doubles per lane: accumulators + the 2M frame values
#endif
// entries per part / parts per team for a given (M, KC)
__host__ __device__ constexpr int cov_nep_max(int M, int KC) {
    int n = (OIVA_COV_REG_BUDGET - 2 * M) / (2 * KC);
    return n < 1 ? 1 : n;
}
__host__ __device__ constexpr int cov_parts(int M, int KC) {
    return (oiva_tri(M) + cov_nep_max(M, KC) - 1) / cov_nep_max(M, KC);
}
// frames per stage (even): ~16 KB of fp64 samples for M <= 8 (many single-warp teams per SM), ~32 KB for the
// blocked many-channel kernel (one big team per SM: fewer, larger hand-offs)
__host__ __device__ constexpr int cov_chunk_frames(int M) {
    int tc = (M <= 8 ? 32 : 64) / M;
    tc &= ~1;
    return tc < 2 ? 2 : (tc > 16 ? 16 : tc);
}
__host__ __device__ constexpr int cov_teams_per_cta(int P) { return (8 / P) < 1 ? 1 : (8 / P); }
__host__ __device__ constexpr int cov_threads(int P) { return cov_teams_per_cta(P) * P * 32; }

struct CovParams {
    const void* Xg;      // grouped samples
    const double* phi;   // (B, K, Tp)
    cplx* Vg;            // (G, K, NE, 32)
    cplx* Vpart;         // (nsplit, G, K, NE, 32) per-split partial sums (deterministic mode), or nullptr: atomics into Vg
    GroupLayout L;
    long long G;         // groups in total = B * NG
    int NGphi;           // groups per phi block (NG, or G when one block of ones serves every group)
    int K, k0;
    int nsplit;          // >1: the frames of a group are split over several teams; <= 0 on entry of the launcher: choose
    int max_split;       // largest split count the scratch buffer has slots for
    int stages;          // ring depth S
    double invT;
};

template <typename ST, int M, int KC, int P, int PART>
struct CovPart {
    typedef typename StoreC<ST>::type XC;
    static constexpr int NE = oiva_tri(M);
    static constexpr int NEP = (NE + P - 1) / P;
    static constexpr int NACC = NEP;
    static constexpr int TC = cov_chunk_frames(M);

    // accumulate `nfr` (<= TC) frames of a staged chunk: xs = [TC][M][32] complex, ph = [KC][TC]
    __device__ static __forceinline__ void accumulate(cplx (&acc)[NEP][KC], const XC* __restrict__ xs,
                                                      const double* __restrict__ ph, int nfr, int lane, int) {
#pragma unroll
        for (int fr = 0; fr < TC; ++fr) {
            if (fr < nfr) {
                cplx x[M];
                double w[KC];
#pragma unroll
                for (int c = 0; c < M; ++c) x[c] = widen(xs[(fr * M + c) * OIVA_GROUP + lane]);
#pragma unroll
                for (int k = 0; k < KC; ++k) w[k] = ph[k * TC + fr];
                static_for<NEP>([&](auto nc) {
                    constexpr int n = decltype(nc)::value;
                    constexpr int e = PART + n * P;
                    if constexpr (e < NE) {
                        constexpr int i = ent_row(e), j = ent_col(e);
                        if constexpr (i == j) {
                            const double pr = fma(x[i].x, x[i].x, x[i].y * x[i].y);
#pragma unroll
                            for (int k = 0; k < KC; ++k) acc[n][k].x = fma(w[k], pr, acc[n][k].x);
                        } else {
                            const double pr = fma(x[i].x, x[j].x, x[i].y * x[j].y);
                            const double pi = fma(x[i].y, x[j].x, -(x[i].x * x[j].y));
#pragma unroll
                            for (int k = 0; k < KC; ++k) {
                                acc[n][k].x = fma(w[k], pr, acc[n][k].x);
                                acc[n][k].y = fma(w[k], pi, acc[n][k].y);
                            }
                        }
                    }
                });
            }
        }
    }

    // write this lane's entries of the group
    __device__ static __forceinline__ void finish(const cplx (&acc)[NEP][KC], cplx* __restrict__ Vgrp /* (K,NE,32) */,
                                                  int k0, int K, double invT, bool atomic, int lane, int) {
        static_for<NEP>([&](auto nc) {
            constexpr int n = decltype(nc)::value;
            constexpr int e = PART + n * P;
            if constexpr (e < NE) {
                constexpr bool diag = ent_row(e) == ent_col(e);
#pragma unroll
                for (int k = 0; k < KC; ++k) {
                    if (k0 + k < K) {
                        const cplx v = cmake(acc[n][k].x * invT, diag ? 0.0 : acc[n][k].y * invT);
                        cplx* dst = Vgrp + ((size_t)(k0 + k) * NE + e) * OIVA_GROUP + lane;
                        if (atomic) {
                            atomicAdd(&dst->x, v.x);
                            if (!diag) atomicAdd(&dst->y, v.y);
                        } else {
                            *dst = v;
                        }
                    }
                }
            }
        });
    }
};

// The body run by one warp of a team for its compile-time part.
// Units of work: (group, frame split); a team owns the contiguous unit range [u_begin, u_end).  All
// producer / consumer cursors are advanced incrementally (no divisions in the per-chunk path).
template <typename CP, typename ST, int M, int KC>
__device__ __forceinline__ void cov_team_body(const CovParams& p, unsigned char* team_smem, long long team_global,
                                              long long n_teams_total, int lane, bool is_leader_warp, int part) {
    typedef typename CP::XC XC;
    constexpr int TC = CP::TC;
    const GroupLayout& L = p.L;
    const int S = p.stages;
    const int Tp = L.frame_pitch();
    constexpr size_t x_stage = (size_t)TC * M * OIVA_GROUP * sizeof(XC);
    constexpr size_t stage_bytes = ((x_stage + (size_t)KC * TC * sizeof(double) + 127) / 128) * 128;
    uint64_t* full = reinterpret_cast<uint64_t*>(team_smem);
    uint64_t* empty = full + S;
    unsigned char* stage0 = team_smem + 128 * ((2 * S * sizeof(uint64_t) + 127) / 128);

    const int nchunks = (L.T + TC - 1) / TC;
    const int nsplit = p.nsplit;
    const long long U = p.G * nsplit;
    const int u_begin = (int)(U * team_global / n_teams_total);
    const int u_end = (int)(U * (team_global + 1) / n_teams_total);
    const XC* Xg = reinterpret_cast<const XC*>(p.Xg);
    const size_t frame_elems = L.frame_elems(), group_elems = L.group_elems();
    const bool leader = is_leader_warp && lane == 0;

    // ---- producer state (leader lane only) -------------------------------------------------------
    int pu = u_begin, pc = 0, pce = 0, pstage = 0, puse = 0;
    const XC* psrc = nullptr;
    const double* pphi = nullptr;
    auto producer_unit = [&]() {  // position the producer on unit pu (once per unit)
        const int gi = nsplit == 1 ? pu : pu / nsplit;
        const int sp = pu - gi * nsplit;
        pc = nsplit == 1 ? 0 : (int)((long long)nchunks * sp / nsplit);
        pce = nsplit == 1 ? nchunks : (int)((long long)nchunks * (sp + 1) / nsplit);
        const int b = gi / p.NGphi;
        psrc = Xg + (size_t)gi * group_elems + (size_t)pc * TC * frame_elems;
        pphi = p.phi + (size_t)b * p.K * Tp + (size_t)pc * TC;
    };
    auto issue = [&]() {
        while (pc >= pce) {
            if (++pu >= u_end) return;
            producer_unit();
        }
        if (puse > 0) mbar_wait(&empty[pstage], (puse - 1) & 1);
        const int nfr = min(TC, L.T - pc * TC);
        unsigned char* dst = stage0 + (size_t)pstage * stage_bytes;
        const uint32_t xb = (uint32_t)(nfr * frame_elems * sizeof(XC));
        const uint32_t pb = (uint32_t)(((nfr + 1) & ~1) * sizeof(double));
        mbar_arrive_expect_tx(&full[pstage], xb + KC * pb);
        tma_load_1d(dst, psrc, xb, &full[pstage]);
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            const int ks = min(p.k0 + k, p.K - 1);  // padded source slots re-read the last row (never written back)
            tma_load_1d(dst + x_stage + (size_t)k * TC * sizeof(double), pphi + (size_t)ks * Tp, pb, &full[pstage]);
        }
        psrc += (size_t)TC * frame_elems;
        pphi += TC;
        ++pc;
        if (++pstage == S) {
            pstage = 0;
            ++puse;
        }
    };
    if (leader && pu < u_end) {
        producer_unit();
        for (int i = 0; i < S - 1; ++i) issue();
    } else {
        pu = u_end;  // non-leaders never issue
    }

    // ---- consumers ---------------------------------------------------------------------------------
    int cstage = 0, cphase = 0;
    for (int u = u_begin; u < u_end; ++u) {
        const int gi = nsplit == 1 ? u : u / nsplit;
        const int sp = u - gi * nsplit;
        const int c0 = nsplit == 1 ? 0 : (int)((long long)nchunks * sp / nsplit);
        const int c1 = nsplit == 1 ? nchunks : (int)((long long)nchunks * (sp + 1) / nsplit);
        if (c0 >= c1 && !(nsplit > 1 && p.Vpart)) continue;  // (a partial buffer slot must always be written)
        cplx acc[CP::NACC][KC];
#pragma unroll
        for (int n = 0; n < CP::NACC; ++n)
#pragma unroll
            for (int k = 0; k < KC; ++k) acc[n][k] = cmake(0.0, 0.0);
        for (int c = c0; c < c1; ++c) {
            const int nfr = min(TC, L.T - c * TC);
            if (leader && pu < u_end) issue();
            mbar_wait(&full[cstage], cphase);
            const unsigned char* src = stage0 + (size_t)cstage * stage_bytes;
            CP::accumulate(acc, reinterpret_cast<const XC*>(src), reinterpret_cast<const double*>(src + x_stage), nfr, lane,
                           part);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[cstage]);
            if (++cstage == S) {
                cstage = 0;
                cphase ^= 1;
            }
        }
        const size_t grp_elems = (size_t)p.K * CP::NE * OIVA_GROUP;
        if (nsplit > 1 && p.Vpart)  // this split's own slot, summed in a fixed order by k_cov_sum_partials
            CP::finish(acc, p.Vpart + ((size_t)sp * p.G + gi) * grp_elems, p.k0, p.K, p.invT, false, lane, part);
        else
            CP::finish(acc, p.Vg + (size_t)gi * grp_elems, p.k0, p.K, p.invT, nsplit > 1, lane, part);
    }
}

template <typename ST, int M, int KC, int P, int PART = 0>
struct CovDispatch {
    __device__ static __forceinline__ void run(int part, const CovParams& p, unsigned char* team_smem,
                                               long long team_global, long long n_teams_total, int lane) {
        if (part == PART)
            cov_team_body<CovPart<ST, M, KC, P, PART>, ST, M, KC>(p, team_smem, team_global, n_teams_total, lane, PART == 0,
                                                                 PART);
        else if constexpr (PART + 1 < P)
            CovDispatch<ST, M, KC, P, PART + 1>::run(part, p, team_smem, team_global, n_teams_total, lane);
    }
};

// blockDim.x = teams_per_cta * P * 32; dynamic smem = teams_per_cta * team_smem_bytes
template <typename ST, int M, int KC, int P>
__global__ void __launch_bounds__(cov_threads(P)) k_cov(const CovParams p, int teams_per_cta, int team_smem_bytes) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int team = warp / P, part = warp - team * P;
    unsigned char* team_smem = smem_raw + (size_t)team * team_smem_bytes;
    if (part == 0 && lane == 0) {
        uint64_t* full = reinterpret_cast<uint64_t*>(team_smem);
        uint64_t* empty = full + p.stages;
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], P);
        }
        mbar_fence_init();
    }
    __syncthreads();
    CovDispatch<ST, M, KC, P>::run(part, p, team_smem, (long long)blockIdx.x * teams_per_cta + team,
                                            (long long)gridDim.x * teams_per_cta, lane);
}

// ---------------------------------------------------------------------------------------------------------
// Block-partitioned variant for many channels (M >= 9).  With the 1-D "entry e -> part e % P" split above every
// part is its own unrolled code path; at M = 16 that is 13 paths and the kernel starves on instruction fetch
// (ncu: stall_no_instruction 30 cycles per issue, 8 % fp64 utilisation).  Here the lower triangle is cut into
// BT x BT blocks, a warp's block coordinates (bi, bj) are RUNTIME values (only shared-memory addresses depend on
// them) while the indexing inside a block stays compile-time, so all warps of a team run ONE code path; a block
// needs 2*BT channel loads for BT*BT entries (classic register tiling).  Diagonal blocks compute the full block
// and store only i >= j.  KC <= 2 keeps BT*BT*KC complex accumulators in registers; more sources = more passes.
// ---------------------------------------------------------------------------------------------------------
constexpr int COV_BT = 4;
__host__ __device__ constexpr int cov_blocks_per_dim(int M) { return (M + COV_BT - 1) / COV_BT; }
__host__ __device__ constexpr int cov_block_parts(int M) { return cov_blocks_per_dim(M) * (cov_blocks_per_dim(M) + 1) / 2; }

template <typename ST, int M, int KC>
struct CovBlock {
    typedef typename StoreC<ST>::type XC;
    static constexpr int NE = oiva_tri(M);
    static constexpr int NACC = COV_BT * COV_BT;
    static constexpr int TC = cov_chunk_frames(M);

    __device__ static __forceinline__ void coords(int part, int& bi, int& bj) {
        bi = 0;
        while ((bi + 1) * (bi + 2) / 2 <= part) ++bi;
        bj = part - bi * (bi + 1) / 2;
    }

    __device__ static __forceinline__ void accumulate(cplx (&acc)[NACC][KC], const XC* __restrict__ xs,
                                                      const double* __restrict__ ph, int nfr, int lane, int part) {
        int bi, bj;
        coords(part, bi, bj);
        const XC* xi0 = xs + (size_t)(bi * COV_BT) * OIVA_GROUP + lane;
        const XC* xj0 = xs + (size_t)(bj * COV_BT) * OIVA_GROUP + lane;
#pragma unroll
        for (int fr = 0; fr < TC; ++fr) {
            if (fr < nfr) {
                cplx xi[COV_BT], xj[COV_BT];
                double w[KC];
#pragma unroll
                for (int a = 0; a < COV_BT; ++a) {
                    // channels beyond M (M not a multiple of BT) contribute zeros
                    xi[a] = (bi * COV_BT + a < M) ? widen(xi0[(fr * M + a) * OIVA_GROUP]) : cmake(0.0, 0.0);
                    xj[a] = (bj * COV_BT + a < M) ? widen(xj0[(fr * M + a) * OIVA_GROUP]) : cmake(0.0, 0.0);
                }
#pragma unroll
                for (int k = 0; k < KC; ++k) w[k] = ph[k * TC + fr];
#pragma unroll
                for (int a = 0; a < COV_BT; ++a)
#pragma unroll
                    for (int b = 0; b < COV_BT; ++b) {
                        const double pr = fma(xi[a].x, xj[b].x, xi[a].y * xj[b].y);
                        const double pi = fma(xi[a].y, xj[b].x, -(xi[a].x * xj[b].y));
#pragma unroll
                        for (int k = 0; k < KC; ++k) {
                            acc[a * COV_BT + b][k].x = fma(w[k], pr, acc[a * COV_BT + b][k].x);
                            acc[a * COV_BT + b][k].y = fma(w[k], pi, acc[a * COV_BT + b][k].y);
                        }
                    }
            }
        }
    }

    __device__ static __forceinline__ void finish(const cplx (&acc)[NACC][KC], cplx* __restrict__ Vgrp, int k0, int K,
                                                  double invT, bool atomic, int lane, int part) {
        int bi, bj;
        coords(part, bi, bj);
#pragma unroll
        for (int a = 0; a < COV_BT; ++a)
#pragma unroll
            for (int b = 0; b < COV_BT; ++b) {
                const int i = bi * COV_BT + a, j = bj * COV_BT + b;
                if (i < M && j <= i) {
                    const int e = i * (i + 1) / 2 + j;
#pragma unroll
                    for (int k = 0; k < KC; ++k) {
                        if (k0 + k < K) {
                            const cplx v = cmake(acc[a * COV_BT + b][k].x * invT, i == j ? 0.0 : acc[a * COV_BT + b][k].y * invT);
                            cplx* dst = Vgrp + ((size_t)(k0 + k) * NE + e) * OIVA_GROUP + lane;
                            if (atomic) {
                                atomicAdd(&dst->x, v.x);
                                if (i != j) atomicAdd(&dst->y, v.y);
                            } else {
                                *dst = v;
                            }
                        }
                    }
                }
            }
    }
};

// one team of cov_block_parts(M) warps per CTA.  Register budget: the 10 warps of M = 13..16 put 3 warps on one SM
// sub-partition, 3 x 32 x R <= 16384 caps R at 168 (what ptxas picks from the launch bounds; a larger __maxnreg__ makes
// the launch fail), and 128 of those hold the 4 x 4 x KC=2 complex accumulators -- staging the products of a block row
// before accumulating them (more ILP against the "wait" stalls ncu shows) spills.  The kernel is register-file-bound.
template <typename ST, int M, int KC>
__global__ void __launch_bounds__(cov_block_parts(M) * 32) k_cov_blocked(const CovParams p, int team_smem_bytes) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int P = cov_block_parts(M);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
        uint64_t* empty = full + p.stages;
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], P);
        }
        mbar_fence_init();
    }
    __syncthreads();
    cov_team_body<CovBlock<ST, M, KC>, ST, M, KC>(p, smem_raw, blockIdx.x, gridDim.x, lane, warp == 0, warp);
}

}  // namespace oiva
