// Weighted covariance  V_k[row] = (1/T) sum_t phi_k(t) x(t) x(t)^H  for all K sources in ONE pass over X.
// (reference: overiva.py:179 -- K separate zgemm calls, each streaming X and an |X|-sized temporary;
//  overiva.py:87 for the unweighted input covariance.)
//
// Mapping (DESIGN.md "covariance kernel"):
//  * a "team" of P warps owns a row (one (mixture, bin) pair) at a time; rows arrive in shared memory as
//    whole tiles through a ring of S stages filled by 1-D bulk TMA (cp.async.bulk + mbarrier
//    complete_tx), issued S-1 tiles ahead by the team's leader lane; consumers release a stage through
//    an "empty" mbarrier.  The phi weights of the tile ride in the same stage.
//  * inside a warp, lane <-> frame: each lane reads the 2M planes of its frame (conflict-free 256-byte
//    LDS.64 segments), forms the Hermitian products p_ij = x_i conj(x_j) ONCE for all sources and
//    accumulates phi_k * p_ij for its share of the lower-triangle entries in registers
//    (entry e of the lower triangle belongs to part e % P; a warp's part is compile-time, so all
//    register indexing is static).
//  * after the last tile of the row, the per-lane partial sums are combined with a transposing
//    butterfly (each exchange halves the number of live values, ~N shuffles for N accumulators instead
//    of 5N), scaled by 1/T and written to both triangles of V.
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

__host__ __device__ constexpr int tri(int M) { return M * (M + 1) / 2; }
__host__ __device__ constexpr int ent_row(int e) {
    int i = 0;
    while ((i + 1) * (i + 2) / 2 <= e) ++i;
    return i;
}
__host__ __device__ constexpr int ent_col(int e) { return e - ent_row(e) * (ent_row(e) + 1) / 2; }

#ifndef OIVA_COV_REG_BUDGET
#define OIVA_COV_REG_BUDGET 96  // doubles per lane: accumulators + the 2M frame values
#endif
// entries per part / parts per team for a given (M, KC)
__host__ __device__ constexpr int cov_nep_max(int M, int KC) {
    int n = (OIVA_COV_REG_BUDGET - 2 * M) / (2 * KC);
    return n < 1 ? 1 : n;
}
__host__ __device__ constexpr int cov_parts(int M, int KC) { return (tri(M) + cov_nep_max(M, KC) - 1) / cov_nep_max(M, KC); }
__host__ __device__ constexpr int cov_nep(int M, int KC) { return (tri(M) + cov_parts(M, KC) - 1) / cov_parts(M, KC); }

struct CovParams {
    const void* Xp;      // planar rows
    const double* phi;   // (B, K, Tp)
    double* V;           // (R, K, M, M) complex as doubles
    RowLayout L;
    int R, F, K, k0;
    int nsplit;          // >1: several teams share a row (split over tiles), results added atomically
    int stages;          // ring depth S
    int xpitch;          // frames per plane reserved in a stage (TT, or TL when nT == 1)
    int ppitch;          // phi doubles per source reserved in a stage
    double invT;
};

// transposing butterfly: v[0..N32) per lane -> lane l ends with the totals of indices [Q*l, Q*l + Q)
template <int N32>
__device__ __forceinline__ void xreduce(double (&v)[N32], int lane) {
    static_assert(N32 % 32 == 0, "N32 must be a multiple of 32");
#pragma unroll
    for (int lvl = 0; lvl < 5; ++lvl) {
        const int H = N32 >> (lvl + 1);
        const int off = 16 >> lvl;
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int n = 0; n < H; ++n) {
            double lo = v[n], hi = v[n + H];
            double send = up ? lo : hi;
            double keep = up ? hi : lo;
            v[n] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
}

template <typename ST, int M, int KC, int P, int PART>
struct CovPart {
    static constexpr int NE = tri(M);
    static constexpr int NEP = (NE + P - 1) / P;
    static constexpr int N = NEP * KC * 2;
    static constexpr int N32 = ((N + 31) / 32) * 32;
    static constexpr int Q = N32 / 32;

    // accumulate one tile: xs = [2M][TT] planes (shared or global), ph = [KC][phiN]
    __device__ static __forceinline__ void accumulate(double (&acc)[N32], const ST* __restrict__ xs,
                                                      const double* __restrict__ ph, int TT, int phiN, int nvalid,
                                                      int lane) {
        for (int tl = lane; tl < nvalid; tl += 32) {
            double xr[M], xi[M], w[KC];
#pragma unroll
            for (int c = 0; c < M; ++c) {
                xr[c] = (double)xs[(2 * c) * TT + tl];
                xi[c] = (double)xs[(2 * c + 1) * TT + tl];
            }
#pragma unroll
            for (int k = 0; k < KC; ++k) w[k] = ph[k * phiN + tl];
            static_for<NEP>([&](auto nc) {
                constexpr int n = decltype(nc)::value;
                constexpr int e = PART + n * P;
                if constexpr (e < NE) {
                    constexpr int i = ent_row(e), j = ent_col(e);
                    if constexpr (i == j) {
                        double pr = fma(xr[i], xr[i], xi[i] * xi[i]);
#pragma unroll
                        for (int k = 0; k < KC; ++k) acc[(n * KC + k) * 2] = fma(w[k], pr, acc[(n * KC + k) * 2]);
                    } else {
                        double pr = fma(xr[i], xr[j], xi[i] * xi[j]);
                        double pi = fma(xi[i], xr[j], -(xr[i] * xi[j]));
#pragma unroll
                        for (int k = 0; k < KC; ++k) {
                            acc[(n * KC + k) * 2] = fma(w[k], pr, acc[(n * KC + k) * 2]);
                            acc[(n * KC + k) * 2 + 1] = fma(w[k], pi, acc[(n * KC + k) * 2 + 1]);
                        }
                    }
                }
            });
        }
    }

    // combine the 32 lanes and write the row's entries (both triangles)
    __device__ static __forceinline__ void finish(double (&acc)[N32], double* __restrict__ Vrow /* (K,M,M,2) */,
                                                  int k0, double invT, bool atomic, int lane) {
        xreduce<N32>(acc, lane);
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const int idx = Q * lane + q;
            if (idx < N) {
                const int ri = idx & 1;
                const int k = (idx >> 1) % KC;
                const int n = idx / (2 * KC);
                const int e = PART + n * P;
                if (e < NE) {
                    int i = 0;
                    while ((i + 1) * (i + 2) / 2 <= e) ++i;
                    const int j = e - i * (i + 1) / 2;
                    const double val = acc[q] * invT;
                    double* lo = Vrow + (((size_t)(k0 + k) * M + i) * M + j) * 2 + ri;
                    double* up = Vrow + (((size_t)(k0 + k) * M + j) * M + i) * 2 + ri;
                    if (i == j) {
                        if (ri == 0) {
                            if (atomic) atomicAdd(lo, val); else *lo = val;
                        } else if (!atomic) {
                            *lo = 0.0;
                        }
                    } else {
                        const double uval = ri ? -val : val;
                        if (atomic) {
                            atomicAdd(lo, val);  // upper triangle mirrored afterwards (k_cov_mirror)
                        } else {
                            *lo = val;
                            *up = uval;
                        }
                    }
                }
            }
        }
    }
};

// The body run by one warp of a team for its compile-time part.
template <typename ST, int M, int KC, int P, int PART, bool USE_TMA>
__device__ __forceinline__ void cov_team_body(const CovParams& p, unsigned char* team_smem, int team_global,
                                              int n_teams_total, int lane, bool is_leader_warp) {
    using CP = CovPart<ST, M, KC, P, PART>;
    const RowLayout& L = p.L;
    const int S = p.stages;
    const int Tp = L.frame_pitch();
    const size_t x_stage = (size_t)2 * M * p.xpitch * sizeof(ST);
    const size_t stage_bytes = ((x_stage + (size_t)KC * p.ppitch * sizeof(double) + 127) / 128) * 128;
    uint64_t* full = reinterpret_cast<uint64_t*>(team_smem);
    uint64_t* empty = full + S;
    unsigned char* stage0 = team_smem + 128 * ((2 * S * sizeof(uint64_t) + 127) / 128);

    const long long U = (long long)p.R * p.nsplit;
    const long long u_begin = U * team_global / n_teams_total;
    const long long u_end = U * (team_global + 1) / n_teams_total;
    const size_t row_elems = L.row_elems();
    const ST* Xp = reinterpret_cast<const ST*>(p.Xp);
    const bool leader = is_leader_warp && lane == 0;

    // producer cursor (leader lane only)
    long long pu = u_begin;
    int pt = 0, pte = 0;
    if (pu < u_end) {
        int sp = (int)(pu % p.nsplit);
        pt = (int)((long long)L.nT * sp / p.nsplit);
        pte = (int)((long long)L.nT * (sp + 1) / p.nsplit);
    }
    int q_load = 0;
    auto issue = [&]() {
        while (pu < u_end && pt >= pte) {  // next unit (skipping empty tile ranges)
            ++pu;
            if (pu < u_end) {
                int sp = (int)(pu % p.nsplit);
                pt = (int)((long long)L.nT * sp / p.nsplit);
                pte = (int)((long long)L.nT * (sp + 1) / p.nsplit);
            }
        }
        if (pu >= u_end) return;
        const int stage = q_load % S, use = q_load / S;
        if (use > 0) mbar_wait(&empty[stage], (use - 1) & 1);
        const long long row = pu / p.nsplit;
        const int b = (int)(row / p.F);
        unsigned char* dst = stage0 + (size_t)stage * stage_bytes;
        const uint32_t xb = (uint32_t)((size_t)2 * M * L.pitch(pt) * sizeof(ST));
        int pc = (L.valid(pt) + 1) & ~1;  // phi doubles to copy (16-byte multiple)
        const uint32_t pb = (uint32_t)(pc * sizeof(double));
        mbar_arrive_expect_tx(&full[stage], xb + KC * pb);
        tma_load_1d(dst, Xp + (size_t)row * row_elems + L.tile_off(pt), xb, &full[stage]);
        const double* ph = p.phi + ((size_t)b * p.K + p.k0) * Tp + (size_t)pt * L.TT;
#pragma unroll
        for (int k = 0; k < KC; ++k)
            tma_load_1d(dst + x_stage + (size_t)k * p.ppitch * sizeof(double), ph + (size_t)k * Tp, pb, &full[stage]);
        ++q_load;
        ++pt;
    };

    if (USE_TMA && leader) {
        for (int i = 0; i < S - 1; ++i) issue();
    }

    int q_cons = 0;
    for (long long u = u_begin; u < u_end; ++u) {
        const long long row = u / p.nsplit;
        const int sp = (int)(u % p.nsplit);
        const int tb = (int)((long long)L.nT * sp / p.nsplit);
        const int te = (int)((long long)L.nT * (sp + 1) / p.nsplit);
        if (tb >= te) continue;
        double acc[CP::N32];
#pragma unroll
        for (int n = 0; n < CP::N32; ++n) acc[n] = 0.0;
        for (int t = tb; t < te; ++t) {
            const int nvalid = L.valid(t);
            if (USE_TMA) {
                if (leader) issue();
                const int stage = q_cons % S;
                mbar_wait(&full[stage], (q_cons / S) & 1);
                const unsigned char* src = stage0 + (size_t)stage * stage_bytes;
                CP::accumulate(acc, reinterpret_cast<const ST*>(src), reinterpret_cast<const double*>(src + x_stage),
                               L.pitch(t), p.ppitch, nvalid, lane);
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[stage]);
                ++q_cons;
            } else {
                const int b = (int)(row / p.F);
                const double* ph = p.phi + ((size_t)b * p.K + p.k0) * Tp + (size_t)t * L.TT;
                CP::accumulate(acc, Xp + (size_t)row * row_elems + L.tile_off(t), ph, L.pitch(t), Tp, nvalid, lane);
            }
        }
        CP::finish(acc, p.V + (size_t)row * p.K * M * M * 2, p.k0, p.invT, p.nsplit > 1, lane);
    }
}

template <typename ST, int M, int KC, int P, bool USE_TMA, int PART = 0>
struct CovDispatch {
    __device__ static __forceinline__ void run(int part, const CovParams& p, unsigned char* team_smem, int team_global,
                                               int n_teams_total, int lane) {
        if (part == PART)
            cov_team_body<ST, M, KC, P, PART, USE_TMA>(p, team_smem, team_global, n_teams_total, lane, PART == 0);
        else if constexpr (PART + 1 < P)
            CovDispatch<ST, M, KC, P, USE_TMA, PART + 1>::run(part, p, team_smem, team_global, n_teams_total, lane);
    }
};

__host__ __device__ constexpr int cov_teams_per_cta(int P) { return (8 / P) < 1 ? 1 : (8 / P); }
__host__ __device__ constexpr int cov_threads(int P) { return cov_teams_per_cta(P) * P * 32; }

// blockDim.x = n_teams * P * 32; dynamic smem = n_teams * team_smem_bytes
template <typename ST, int M, int KC, int P, bool USE_TMA>
__global__ void __launch_bounds__(cov_threads(P)) k_cov(const CovParams p, int teams_per_cta, int team_smem_bytes) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int team = warp / P, part = warp - team * P;
    unsigned char* team_smem = smem_raw + (size_t)team * team_smem_bytes;
    if (USE_TMA) {
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
    }
    CovDispatch<ST, M, KC, P, USE_TMA>::run(part, p, team_smem, blockIdx.x * teams_per_cta + team,
                                            gridDim.x * teams_per_cta, lane);
}

// split-row path only: copy the (atomically accumulated) lower triangle into the upper one so that V is
// exactly Hermitian; one thread per (row, k, i, j) with i > j
__global__ void k_cov_mirror(double* __restrict__ V, long long R, int K, int k0, int KC, int M);

}  // namespace oiva
