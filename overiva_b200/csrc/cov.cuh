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
// (3 parts would leave 6 of the 8 warps of a CTA running -- two 3-warp teams, two of the four schedulers with a single
// warp: ncu showed the fp64 pipe 39 % busy at M = K = 6 -- so 3 is rounded up to 4: two 4-warp teams)
__host__ __device__ constexpr int cov_parts(int M, int KC) {
    return (oiva_tri(M) + cov_nep_max(M, KC) - 1) / cov_nep_max(M, KC) == 3
               ? 4
               : (oiva_tri(M) + cov_nep_max(M, KC) - 1) / cov_nep_max(M, KC);
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
    // fused covariance + IP sweep (cov_sweep.cuh): the loop state the epilogue of a group updates.  Unused otherwise.
    cplx* Wg;              // (G, M*M, 32) grouped W_hat
    const cplx* Cg;        // (G, NE, 32) grouped input covariance
    const double* wscale;  // (B, K) or nullptr
    int* status;           // (B) per-mixture status words
};

// epilogue of a group in cov_team_body: by default the lane's entries go to Vg (CovPart::finish); the fused kernel
// passes a policy whose run() consumes the accumulators in registers instead
struct CovStoreEpilogue {};

// WBIN = false: one weight per (source, frame), phi (B, K, Tp) -- the IVA source models (overiva.py:152-173).
// WBIN = true : one weight per (source, frame, BIN), a grouped array [gi][k][Tp][32] -- ILRMA's low-rank source model
//               (1 / r_k(f, t)); the weights of a chunk are staged per lane next to the samples.
template <typename ST, int M, int KC, int P, int PART, bool WBIN = false>
struct CovPart {
    typedef typename StoreC<ST>::type XC;
    static constexpr int NE = oiva_tri(M);
    static constexpr int NEP = (NE + P - 1) / P;
    static constexpr int NACC = NEP;
    // frames per ring stage.  (Stages twice as long for the four-warp teams of the many-source shapes -- half as many
    // hand-offs -- measured much slower: M = K = 6 2.35 -> 3.67 ms, M = K = 8 5.5 -> 9.1 ms per 256 mixtures.)
    static constexpr int TC = cov_chunk_frames(M);
    static constexpr bool WB = WBIN;
    static constexpr int WLANES = WBIN ? OIVA_GROUP : 1;  // weights per (source, frame) in a stage
    // the branch-free whole-chunk path (below) pays up to 6 channels (bench shape 2.21 -> 2.10 ms per launch, M = K = 4
    // 0.77 -> 0.71, M = K = 6 2.39 -> 2.35); with 8 channel values live per frame it costs (M = K = 8: 5.53 -> 6.16 ms)
    static constexpr bool USE_WHOLE = M <= 6;

    // accumulate `nfr` (<= TC) frames of a staged chunk: xs = [TC][M][32] complex, ph = [KC][TC] (or [KC][TC][32]).
    // WHOLE: a complete chunk -- no per-frame test, so the loads of a frame can be scheduled over the arithmetic of the
    // previous one (every chunk but a group's last is complete).
    template <bool WHOLE = false>
    __device__ static __forceinline__ void accumulate(cplx (&acc)[NEP][KC], const XC* __restrict__ xs,
                                                      const double* __restrict__ ph, int nfr, int lane, int) {
#pragma unroll
        for (int fr = 0; fr < TC; ++fr) {
            if (WHOLE || fr < nfr) {
                cplx x[M];
                double w[KC];
#pragma unroll
                for (int c = 0; c < M; ++c) x[c] = widen(xs[(fr * M + c) * OIVA_GROUP + lane]);
#pragma unroll
                for (int k = 0; k < KC; ++k) w[k] = WBIN ? ph[(k * TC + fr) * OIVA_GROUP + lane] : ph[k * TC + fr];
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
template <typename CP, typename ST, int M, int KC, typename EP = CovStoreEpilogue>
__device__ __forceinline__ void cov_team_body(const CovParams& p, unsigned char* team_smem, long long team_global,
                                              long long n_teams_total, int lane, bool is_leader_warp, int part) {
    typedef typename CP::XC XC;
    constexpr int TC = CP::TC;
    const GroupLayout& L = p.L;
    const int S = p.stages;
    const int Tp = L.frame_pitch();
    constexpr size_t x_stage = (size_t)TC * M * OIVA_GROUP * sizeof(XC);
    constexpr int WL = CP::WLANES;
    constexpr size_t stage_bytes = ((x_stage + (size_t)KC * TC * WL * sizeof(double) + 127) / 128) * 128;
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
        psrc = Xg + (size_t)gi * group_elems + (size_t)pc * TC * frame_elems;
        if constexpr (CP::WB) {  // per-bin weights: [gi][k][Tp][32]
            pphi = p.phi + ((size_t)gi * p.K * Tp + (size_t)pc * TC) * OIVA_GROUP;
        } else {
            const int b = gi / p.NGphi;
            pphi = p.phi + (size_t)b * p.K * Tp + (size_t)pc * TC;
        }
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
        const uint32_t pb = (uint32_t)((CP::WB ? nfr * OIVA_GROUP : ((nfr + 1) & ~1)) * sizeof(double));
        mbar_arrive_expect_tx(&full[pstage], xb + KC * pb);
        tma_load_1d(dst, psrc, xb, &full[pstage]);
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            const int ks = min(p.k0 + k, p.K - 1);  // padded source slots re-read the last row (never written back)
            tma_load_1d(dst + x_stage + (size_t)k * TC * WL * sizeof(double), pphi + (size_t)ks * Tp * WL, pb, &full[pstage]);
        }
        psrc += (size_t)TC * frame_elems;
        pphi += TC * WL;
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
            if (CP::USE_WHOLE && nfr == TC)
                CP::template accumulate<true>(acc, reinterpret_cast<const XC*>(src),
                                              reinterpret_cast<const double*>(src + x_stage), nfr, lane, part);
            else
                CP::template accumulate<false>(acc, reinterpret_cast<const XC*>(src),
                                               reinterpret_cast<const double*>(src + x_stage), nfr, lane, part);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[cstage]);
            if (++cstage == S) {
                cstage = 0;
                cphase ^= 1;
            }
        }
        if constexpr (!std::is_same<EP, CovStoreEpilogue>::value) {
            EP::run(acc, p, gi, lane);  // (launched with nsplit == 1 only)
            continue;
        }
        const size_t grp_elems = (size_t)p.K * CP::NE * OIVA_GROUP;
        if (nsplit > 1 && p.Vpart)  // this split's own slot, summed in a fixed order by k_cov_sum_partials
            CP::finish(acc, p.Vpart + ((size_t)sp * p.G + gi) * grp_elems, p.k0, p.K, p.invT, false, lane, part);
        else
            CP::finish(acc, p.Vg + (size_t)gi * grp_elems, p.k0, p.K, p.invT, nsplit > 1, lane, part);
    }
}

template <typename ST, int M, int KC, int P, bool WBIN, int PART = 0>
struct CovDispatch {
    __device__ static __forceinline__ void run(int part, const CovParams& p, unsigned char* team_smem,
                                               long long team_global, long long n_teams_total, int lane) {
        if (part == PART)
            cov_team_body<CovPart<ST, M, KC, P, PART, WBIN>, ST, M, KC>(p, team_smem, team_global, n_teams_total, lane,
                                                                       PART == 0, PART);
        else if constexpr (PART + 1 < P)
            CovDispatch<ST, M, KC, P, WBIN, PART + 1>::run(part, p, team_smem, team_global, n_teams_total, lane);
    }
};

// blockDim.x = teams_per_cta * P * 32; dynamic smem = teams_per_cta * team_smem_bytes
template <typename ST, int M, int KC, int P, bool WBIN = false>
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
    CovDispatch<ST, M, KC, P, WBIN>::run(part, p, team_smem, (long long)blockIdx.x * teams_per_cta + team,
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
    static constexpr bool WB = false;
    static constexpr int WLANES = 1;
    static constexpr bool USE_WHOLE = false;

    __device__ static __forceinline__ void coords(int part, int& bi, int& bj) {
        bi = 0;
        while ((bi + 1) * (bi + 2) / 2 <= part) ++bi;
        bj = part - bi * (bi + 1) / 2;
    }

    template <bool WHOLE = false>
    __device__ static __forceinline__ void accumulate(cplx (&acc)[NACC][KC], const XC* __restrict__ xs,
                                                      const double* __restrict__ ph, int nfr, int lane, int part) {
        int bi, bj;
        coords(part, bi, bj);
        const XC* xi0 = xs + (size_t)(bi * COV_BT) * OIVA_GROUP + lane;
        const XC* xj0 = xs + (size_t)(bj * COV_BT) * OIVA_GROUP + lane;
#pragma unroll
        for (int fr = 0; fr < TC; ++fr) {
            if (WHOLE || fr < nfr) {
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

// ---------------------------------------------------------------------------------------------------------
// Tiled variant for many channels AND several sources (M >= 9, K >= 3; config 5: M = 16, K = 4).
//
// The blocked kernel above keeps 4 x 4 entries x 2 sources per warp, so four sources cost two passes over X and
// every pass re-forms the products x_i conj(x_j): 2 x (4 + 2*2) = 16 fp64-pipe operations per entry and frame.
// Here a warp keeps a 2 x 4 tile for FOUR sources (the same 64 complex accumulators): the product is formed once,
// 4 + 2*4 = 12 operations per entry and frame, and X is streamed ONCE per epoch.  The 4 x 4 diagonal blocks get
// their own tile type -- the 10 entries i >= j, real-only diagonal accumulators -- which costs exactly as much as a
// 2 x 4 tile (6 x 12 + 4 x 6 = 96 operations per frame), so nothing is computed that is not stored (the blocked kernel
// computes 160 entries for the 136 of M = 16).  NB^2 tiles per bin group (NB = ceil(M/4)): NB(NB-1) full + NB diagonal.
// They do not fit the registers of one SM (a tile takes ~200 registers per thread), so the tiles of a group are
// split over the TWO CTAs of a thread-block CLUSTER ("halves", CovTiling<M>::WARPS warps each).  Both halves need the
// same frames: every chunk of frames is fetched from L2 / HBM ONCE -- each CTA issues half of it as a bulk-TMA copy
// with .multicast::cluster, which lands in the shared memory of both CTAs and signals both CTAs' "full" barriers; a stage
// is re-filled when the consumers of BOTH CTAs have released it (they arrive on the local and on the peer's "empty"
// barrier).  8 warps per SM = 2 per scheduler with up to 255 registers each; the per-frame work of every warp is
// identical (96 DFMA-pipe operations), so the four schedulers stay balanced.
// ---------------------------------------------------------------------------------------------------------
template <int M>
struct CovTiling {
    static constexpr int NB = (M + 3) / 4;
    static constexpr int NFULL = NB * (NB - 1);  // 2 x 4 tiles of the strictly-lower 4 x 4 blocks
    static constexpr int NDIAG = NB;             // 10-entry tiles of the diagonal blocks
    static constexpr int FULL_H = NFULL / 2;     // full tiles per half
    static constexpr int DIAG_H0 = (NDIAG + 1) / 2;
    static constexpr int WARPS = FULL_H + DIAG_H0;  // NB = 4: 6 + 2 = 8 warps per CTA; NB = 3: 3 + 2 = 5
    static constexpr int KC = 4;
    static constexpr int TC = 8;                 // frames per stage (M = 16, complex128: 64 KB; 3 stages)
};

// one 2 x 4 tile (rows r0, r0+1; columns c0..c0+3; all below the diagonal blocks) for KC = 4 sources
template <typename ST, int M>
struct CovTileFull {
    typedef typename StoreC<ST>::type XC;
    static constexpr int KC = 4, TC = CovTiling<M>::TC, NE = oiva_tri(M);
    cplx acc[2][4][KC];
    int r0, c0;
    __device__ __forceinline__ void init(int ft) {  // ft: index among the full tiles of the group
        const int blk = ft >> 1;
        int bi = 1;
        while (bi * (bi + 1) / 2 <= blk) ++bi;  // off-diagonal blocks (1,0) (2,0) (2,1) (3,0) ...
        const int bj = blk - bi * (bi - 1) / 2;
        r0 = bi * 4 + (ft & 1) * 2;
        c0 = bj * 4;
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b)
#pragma unroll
                for (int k = 0; k < KC; ++k) acc[a][b][k] = cmake(0.0, 0.0);
    }
    // WHOLE: a complete chunk of TC frames (no per-frame test: the loads of a frame can be hoisted over the arithmetic
    // of the previous one)
    template <bool WHOLE>
    __device__ __forceinline__ void accumulate(const XC* __restrict__ xs, const double* __restrict__ ph, int nfr, int lane) {
        const XC* xr = xs + (size_t)r0 * OIVA_GROUP + lane;
        const XC* xc = xs + (size_t)c0 * OIVA_GROUP + lane;
        const bool r_ok0 = r0 < M, r_ok1 = r0 + 1 < M;  // channels beyond M (M not a multiple of 4) contribute zeros
#pragma unroll
        for (int fr = 0; fr < TC; ++fr) {
            if (WHOLE || fr < nfr) {
                cplx xi[2], xj[4];
                double w[KC];
                xi[0] = r_ok0 ? widen(xr[(fr * M + 0) * OIVA_GROUP]) : cmake(0.0, 0.0);
                xi[1] = r_ok1 ? widen(xr[(fr * M + 1) * OIVA_GROUP]) : cmake(0.0, 0.0);
#pragma unroll
                for (int b = 0; b < 4; ++b) xj[b] = widen(xc[(fr * M + b) * OIVA_GROUP]);  // c0 + 3 < 4*(NB-1) <= M
#pragma unroll
                for (int k = 0; k < KC; ++k) w[k] = ph[k * TC + fr];
#pragma unroll
                for (int a = 0; a < 2; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const double pr = fma(xi[a].x, xj[b].x, xi[a].y * xj[b].y);
                        const double pi = fma(xi[a].y, xj[b].x, -(xi[a].x * xj[b].y));
#pragma unroll
                        for (int k = 0; k < KC; ++k) {
                            acc[a][b][k].x = fma(w[k], pr, acc[a][b][k].x);
                            acc[a][b][k].y = fma(w[k], pi, acc[a][b][k].y);
                        }
                    }
            }
        }
    }
    __device__ __forceinline__ void finish(cplx* __restrict__ Vgrp, int k0, int K, double invT, int lane) const {
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int i = r0 + a, j = c0 + b;
                if (i < M) {
                    const int e = i * (i + 1) / 2 + j;
#pragma unroll
                    for (int k = 0; k < KC; ++k)
                        if (k0 + k < K)
                            Vgrp[((size_t)(k0 + k) * NE + e) * OIVA_GROUP + lane] =
                                cmake(acc[a][b][k].x * invT, acc[a][b][k].y * invT);
                }
            }
    }
};

// the lower triangle (incl. diagonal) of one diagonal 4 x 4 block, KC = 4 sources: 6 complex + 4 real accumulators
template <typename ST, int M>
struct CovTileDiag {
    typedef typename StoreC<ST>::type XC;
    static constexpr int KC = 4, TC = CovTiling<M>::TC, NE = oiva_tri(M);
    cplx off[6][KC];   // (1,0) (2,0) (2,1) (3,0) (3,1) (3,2)
    double dg[4][KC];  // (0,0) (1,1) (2,2) (3,3)
    int d0;
    __device__ __forceinline__ void init(int d) {
        d0 = d * 4;
#pragma unroll
        for (int n = 0; n < 6; ++n)
#pragma unroll
            for (int k = 0; k < KC; ++k) off[n][k] = cmake(0.0, 0.0);
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
            for (int k = 0; k < KC; ++k) dg[n][k] = 0.0;
    }
    template <bool WHOLE>
    __device__ __forceinline__ void accumulate(const XC* __restrict__ xs, const double* __restrict__ ph, int nfr, int lane) {
        const XC* xd = xs + (size_t)d0 * OIVA_GROUP + lane;
#pragma unroll
        for (int fr = 0; fr < TC; ++fr) {
            if (WHOLE || fr < nfr) {
                cplx x[4];
                double w[KC];
#pragma unroll
                for (int a = 0; a < 4; ++a) x[a] = (d0 + a < M) ? widen(xd[(fr * M + a) * OIVA_GROUP]) : cmake(0.0, 0.0);
#pragma unroll
                for (int k = 0; k < KC; ++k) w[k] = ph[k * TC + fr];
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const double pr = fma(x[a].x, x[a].x, x[a].y * x[a].y);
#pragma unroll
                    for (int k = 0; k < KC; ++k) dg[a][k] = fma(w[k], pr, dg[a][k]);
#pragma unroll
                    for (int b = 0; b < a; ++b) {
                        const int n = a * (a - 1) / 2 + b;
                        const double qr = fma(x[a].x, x[b].x, x[a].y * x[b].y);
                        const double qi = fma(x[a].y, x[b].x, -(x[a].x * x[b].y));
#pragma unroll
                        for (int k = 0; k < KC; ++k) {
                            off[n][k].x = fma(w[k], qr, off[n][k].x);
                            off[n][k].y = fma(w[k], qi, off[n][k].y);
                        }
                    }
                }
            }
        }
    }
    __device__ __forceinline__ void finish(cplx* __restrict__ Vgrp, int k0, int K, double invT, int lane) const {
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int i = d0 + a;
            if (i < M) {
#pragma unroll
                for (int k = 0; k < KC; ++k)
                    if (k0 + k < K)
                        Vgrp[((size_t)(k0 + k) * NE + i * (i + 1) / 2 + i) * OIVA_GROUP + lane] = cmake(dg[a][k] * invT, 0.0);
#pragma unroll
                for (int b = 0; b < a; ++b) {
                    const int n = a * (a - 1) / 2 + b;
                    const int e = i * (i + 1) / 2 + d0 + b;
#pragma unroll
                    for (int k = 0; k < KC; ++k)
                        if (k0 + k < K)
                            Vgrp[((size_t)(k0 + k) * NE + e) * OIVA_GROUP + lane] =
                                cmake(off[n][k].x * invT, off[n][k].y * invT);
                }
            }
        }
    }
};

// Consumer loop of one warp over the units (group, frame split) of its cluster; TILE = CovTileFull / CovTileDiag.
// `active` = false: an idle warp (NB = 3: the second half has one tile fewer) that only keeps the barriers moving.
// Lane 0 of warp 0 of EACH CTA is that CTA's producer: it issues its half of every chunk (multicast to both CTAs) and
// the CTA's own phi rows, never blocks (a stage that is still in use is retried later) and keeps polling while its
// warp waits for data, so the ring refills as fast as stages are released.
// Measured alternatives (config 5, one B200, ms per covariance pass; profiles/r02_cov_tiled_variants.md): this version
// 3.89; the producer also polling between the two halves of a chunk through an out-of-line function 4.86 (its warp
// becomes the straggler every stage release waits for); every warp's lane 0 polling behind a shared-memory try-lock
// 4.28-4.43; 4 frames per stage with 4 / 6 / 8 stages 5.4 / 4.8 / 4.5.
template <typename TILE, typename ST, int M>
__device__ __forceinline__ void cov_tiled_consume(const CovParams& p, unsigned char* smem, int tile_index, bool active,
                                                  int lane, bool is_producer_warp, long long pair, long long n_pairs,
                                                  uint32_t rank) {
    typedef typename StoreC<ST>::type XC;
    constexpr int TC = CovTiling<M>::TC, KC = CovTiling<M>::KC, NE = oiva_tri(M);
    const GroupLayout& L = p.L;
    const int S = p.stages;
    const int Tp = L.frame_pitch();
    constexpr size_t x_stage = (size_t)TC * M * OIVA_GROUP * sizeof(XC);
    constexpr size_t stage_bytes = ((x_stage + (size_t)KC * TC * sizeof(double) + 127) / 128) * 128;
    constexpr uint32_t frame_bytes = (uint32_t)(M * OIVA_GROUP * sizeof(XC));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + S;
    unsigned char* stage0 = smem + 128 * ((2 * S * sizeof(uint64_t) + 127) / 128);
    const int nchunks = (L.T + TC - 1) / TC;
    const int nsplit = p.nsplit;
    const long long U = p.G * nsplit;
    const XC* Xg = reinterpret_cast<const XC*>(p.Xg);
    const size_t frame_elems = L.frame_elems(), group_elems = L.group_elems();
    const bool leader = is_producer_warp && lane == 0;
    const uint32_t peer_empty0 = cluster_map(smem_u32(empty), rank ^ 1u);

    // ---- producer state (leader lane) ----------------------------------------------------------------------------
    long long pu = pair;
    int pc = 0, pce = 0, pstage = 0, puse = 0;
    const XC* psrc = nullptr;
    const double* pphi = nullptr;
    auto producer_unit = [&]() {
        const long long gi = pu / nsplit;
        const int sp = (int)(pu - gi * nsplit);
        pc = (int)((long long)nchunks * sp / nsplit);
        pce = (int)((long long)nchunks * (sp + 1) / nsplit);
        const long long b = gi / p.NGphi;
        psrc = Xg + (size_t)gi * group_elems + (size_t)pc * TC * frame_elems;
        pphi = p.phi + (size_t)b * p.K * Tp + (size_t)pc * TC;
    };
    auto try_issue = [&]() {  // issue every chunk whose stage is free (in both CTAs), without blocking
        while (pu < U) {
            if (pc >= pce) {
                pu += n_pairs;
                if (pu < U) producer_unit();
                continue;
            }
            if (puse > 0 && !mbar_test_cluster(&empty[pstage], (puse - 1) & 1)) return;
            const int nfr = min(TC, L.T - pc * TC);
            unsigned char* dst = stage0 + (size_t)pstage * stage_bytes;
            const uint32_t pb = (uint32_t)(((nfr + 1) & ~1) * sizeof(double));
            mbar_arrive_expect_tx(&full[pstage], (uint32_t)nfr * frame_bytes + KC * pb);
            // this CTA's half of the frames of the chunk, delivered to both CTAs
            const int h0 = min(nfr, TC / 2);
            const int f0 = rank ? h0 : 0, fn = rank ? nfr - h0 : h0;
            if (fn > 0)
                tma_load_1d_multicast(dst + (size_t)f0 * frame_bytes, psrc + (size_t)f0 * frame_elems,
                                      (uint32_t)fn * frame_bytes, &full[pstage], (uint16_t)3);
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
        }
    };
    if (leader && pu < U) {
        producer_unit();
        try_issue();
    } else if (!leader) {
        pu = U;
    }

    int cstage = 0, cphase = 0;
    for (long long u = pair; u < U; u += n_pairs) {
        const long long gi = u / nsplit;
        const int sp = (int)(u - gi * nsplit);
        const int c0 = (int)((long long)nchunks * sp / nsplit);
        const int c1 = (int)((long long)nchunks * (sp + 1) / nsplit);
        TILE tile;
        tile.init(tile_index);
        for (int c = c0; c < c1; ++c) {
            const int nfr = min(TC, L.T - c * TC);
            if (leader) {
                while (!mbar_test(&full[cstage], cphase)) try_issue();
                try_issue();
            }
            __syncwarp();
            mbar_wait(&full[cstage], cphase);
            const unsigned char* src = stage0 + (size_t)cstage * stage_bytes;
            if (active) {
                if (nfr == TC)
                    tile.template accumulate<true>(reinterpret_cast<const XC*>(src),
                                                   reinterpret_cast<const double*>(src + x_stage), nfr, lane);
                else
                    tile.template accumulate<false>(reinterpret_cast<const XC*>(src),
                                                    reinterpret_cast<const double*>(src + x_stage), nfr, lane);
            }
            __syncwarp();
            if (lane == 0) {  // release the stage in both CTAs (each producer writes into both)
                mbar_arrive(&empty[cstage]);
                mbar_arrive_cluster(peer_empty0 + (uint32_t)cstage * 8u);
            }
            if (++cstage == S) {
                cstage = 0;
                cphase ^= 1;
            }
        }
        if (active) {
            const size_t grp_elems = (size_t)p.K * NE * OIVA_GROUP;
            cplx* dst = (nsplit > 1) ? p.Vpart + ((size_t)sp * p.G + gi) * grp_elems : p.Vg + (size_t)gi * grp_elems;
            tile.finish(dst, p.k0, p.K, p.invT, lane);
        }
    }
}

// grid = 2 x pairs, cluster dimension 2: cluster j runs the units j, j + pairs, ...; CTA rank h owns half h of the
// tiles.  nsplit > 1 REQUIRES p.Vpart (per-split slots, summed afterwards in a fixed order): no atomics here.
template <typename ST, int M>
__global__ void __launch_bounds__(CovTiling<M>::WARPS * 32, 1) k_cov_tiled(const CovParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    typedef CovTiling<M> TL;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
        uint64_t* empty = full + p.stages;
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full[s], 1);               // the CTA's own producer (expect_tx); bytes come from both CTAs
            mbar_init(&empty[s], 2 * TL::WARPS);  // the consumer warps of BOTH CTAs
        }
        mbar_fence_init();
    }
    cluster_sync_all();  // both CTAs' barriers exist before anyone copies into / arrives on the peer's
    const uint32_t rank = cluster_ctarank();
    const long long pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    if (warp < TL::FULL_H) {
        cov_tiled_consume<CovTileFull<ST, M>, ST, M>(p, smem_raw, (int)rank * TL::FULL_H + warp, true, lane, warp == 0, pair,
                                                    n_pairs, rank);
    } else {
        const int d = (rank ? TL::DIAG_H0 : 0) + (warp - TL::FULL_H);
        cov_tiled_consume<CovTileDiag<ST, M>, ST, M>(p, smem_raw, d, d < TL::NDIAG, lane, false, pair, n_pairs, rank);
    }
    cluster_sync_all();  // nobody leaves while the peer may still signal this CTA's barriers
}

}  // namespace oiva
