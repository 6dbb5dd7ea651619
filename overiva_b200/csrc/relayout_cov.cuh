// Relayout + input covariance in ONE pass over the caller's X (B,T,F,M):
//   Xg[gi][t][c][lane] = X[b][t][f][c]                       (replaces overiva.py:131-132, the swapaxes copy)
//   Cg[gi][e][lane]    = (1/T) sum_t x_i(t) conj(x_j(t))      (replaces overiva.py:87 / ive.py:97 / auxiva_pca.py:71)
// Separately these are a transposing copy (read X, write Xg) and a streaming covariance (read Xg): three passes
// over the samples.  Here a single-warp team (lane <-> bin, as in cov.cuh) pulls the frames of its bin group
// straight from the caller's layout -- for one frame the 32 bins x M channels of a group are ONE contiguous block,
// i.e. one 1-D bulk-TMA copy per frame into a 2-stage shared-memory ring -- and every lane then reads ITS bin's M
// values, writes them to the grouped layout (512-byte coalesced stores per channel) and accumulates the Hermitian
// products in registers: two passes.  The arithmetic and its order are those of k_cov with unit weights, so the
// covariance is bit-identical to the two-kernel path.
#pragma once
#include "cov.cuh"

namespace oiva {

struct RelayoutCovParams {
    const void* X;   // (B, T, F, M) interleaved complex
    void* Xg;        // grouped samples (output)
    cplx* Cg;        // (G, NE, 32) lower-triangle covariance (output)
    cplx* Cpart;     // (nsplit, G, NE, 32) per-split partial sums when nsplit > 1
    GroupLayout L;
    long long G;
    int nsplit;
    int stages;
    double invT;
};

template <typename ST, int M>
__global__ void __launch_bounds__(256) k_relayout_cov(const RelayoutCovParams p, int teams_per_cta, int team_smem_bytes) {
    typedef typename StoreC<ST>::type XC;
    constexpr int NE = oiva_tri(M);
    constexpr int TC = cov_chunk_frames(M);
    constexpr size_t row_bytes_full = (size_t)OIVA_GROUP * M * sizeof(XC);  // one frame of a full group
    constexpr size_t stage_bytes = ((TC * row_bytes_full + 127) / 128) * 128;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int team = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* team_smem = smem_raw + (size_t)team * team_smem_bytes;
    const int S = p.stages;
    uint64_t* full = reinterpret_cast<uint64_t*>(team_smem);
    uint64_t* empty = full + S;
    unsigned char* stage0 = team_smem + 128 * ((2 * S * sizeof(uint64_t) + 127) / 128);
    if (lane == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const GroupLayout& L = p.L;
    const int nchunks = (L.T + TC - 1) / TC;
    const int nsplit = p.nsplit;
    const long long team_global = (long long)blockIdx.x * teams_per_cta + team;
    const long long n_teams = (long long)gridDim.x * teams_per_cta;
    const long long U = p.G * nsplit;
    const int u_begin = (int)(U * team_global / n_teams);
    const int u_end = (int)(U * (team_global + 1) / n_teams);
    const XC* X = reinterpret_cast<const XC*>(p.X);
    XC* Xg = reinterpret_cast<XC*>(p.Xg);
    const bool leader = lane == 0;

    // ---- producer (lane 0): one bulk copy per frame of the chunk ---------------------------------
    int pu = u_begin, pc = 0, pce = 0, pstage = 0, puse = 0;
    size_t p_row0 = 0;     // element offset of (b, t = 0, f0, 0) in X
    uint32_t p_row_bytes = 0;
    auto producer_unit = [&]() {
        const int gi = nsplit == 1 ? pu : pu / nsplit;
        const int sp = pu - gi * nsplit;
        pc = nsplit == 1 ? 0 : (int)((long long)nchunks * sp / nsplit);
        pce = nsplit == 1 ? nchunks : (int)((long long)nchunks * (sp + 1) / nsplit);
        const int b = gi / L.NG, g = gi - b * L.NG;
        const int f0 = g * OIVA_GROUP;
        const int nb = min(OIVA_GROUP, L.F - f0);
        p_row0 = ((size_t)b * L.T * L.F + f0) * M;
        p_row_bytes = (uint32_t)((size_t)nb * M * sizeof(XC));
    };
    auto issue = [&]() {
        while (pc >= pce) {
            if (++pu >= u_end) return;
            producer_unit();
        }
        if (puse > 0) mbar_wait(&empty[pstage], (puse - 1) & 1);
        const int t0 = pc * TC;
        const int nfr = min(TC, L.T - t0);
        unsigned char* dst = stage0 + (size_t)pstage * stage_bytes;
        mbar_arrive_expect_tx(&full[pstage], (uint32_t)nfr * p_row_bytes);
        for (int fr = 0; fr < nfr; ++fr)
            tma_load_1d(dst + (size_t)fr * row_bytes_full, X + p_row0 + (size_t)(t0 + fr) * L.F * M, p_row_bytes,
                        &full[pstage]);
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
        pu = u_end;
    }

    // ---- consumer ---------------------------------------------------------------------------------
    int cstage = 0, cphase = 0;
    for (int u = u_begin; u < u_end; ++u) {
        const int gi = nsplit == 1 ? u : u / nsplit;
        const int sp = u - gi * nsplit;
        const int c0 = nsplit == 1 ? 0 : (int)((long long)nchunks * sp / nsplit);
        const int c1 = nsplit == 1 ? nchunks : (int)((long long)nchunks * (sp + 1) / nsplit);
        const int b = gi / L.NG, g = gi - b * L.NG;
        const bool bin_ok = g * OIVA_GROUP + lane < L.F;
        cplx acc[NE];
#pragma unroll
        for (int e = 0; e < NE; ++e) acc[e] = cmake(0.0, 0.0);
        for (int c = c0; c < c1; ++c) {
            const int t0 = c * TC;
            const int nfr = min(TC, L.T - t0);
            if (leader && pu < u_end) issue();
            mbar_wait(&full[cstage], cphase);
            const XC* xs = reinterpret_cast<const XC*>(stage0 + (size_t)cstage * stage_bytes);
            XC* out = Xg + ((size_t)gi * L.T + t0) * M * OIVA_GROUP + lane;
#pragma unroll
            for (int fr = 0; fr < TC; ++fr) {
                if (fr < nfr) {
                    cplx x[M];
#pragma unroll
                    for (int ch = 0; ch < M; ++ch) {
                        XC v;
                        v.x = 0;
                        v.y = 0;
                        if (bin_ok) v = xs[((size_t)fr * OIVA_GROUP + lane) * M + ch];  // padded bins: zeros
                        out[((size_t)fr * M + ch) * OIVA_GROUP] = v;
                        x[ch] = widen(v);
                    }
                    static_for<NE>([&](auto ec) {
                        constexpr int e = decltype(ec)::value;
                        constexpr int i = ent_row(e), j = ent_col(e);
                        if constexpr (i == j) {
                            acc[e].x = fma(1.0, fma(x[i].x, x[i].x, x[i].y * x[i].y), acc[e].x);
                        } else {
                            const double pr = fma(x[i].x, x[j].x, x[i].y * x[j].y);
                            const double pi = fma(x[i].y, x[j].x, -(x[i].x * x[j].y));
                            acc[e].x = fma(1.0, pr, acc[e].x);
                            acc[e].y = fma(1.0, pi, acc[e].y);
                        }
                    });
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[cstage]);
            if (++cstage == S) {
                cstage = 0;
                cphase ^= 1;
            }
        }
        cplx* dst = (nsplit > 1 ? p.Cpart + ((size_t)sp * p.G + gi) * NE * OIVA_GROUP : p.Cg + (size_t)gi * NE * OIVA_GROUP) + lane;
        static_for<NE>([&](auto ec) {
            constexpr int e = decltype(ec)::value;
            constexpr bool diag = ent_row(e) == ent_col(e);
            dst[(size_t)e * OIVA_GROUP] = cmake(acc[e].x * p.invT, diag ? 0.0 : acc[e].y * p.invT);
        });
        (void)b;
    }
}

}  // namespace oiva
