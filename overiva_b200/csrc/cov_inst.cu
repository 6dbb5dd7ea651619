// Instantiates the covariance kernels for ONE channel count (compiled once per M with -DOIVA_COV_M=<M>,
// so the 16 channel counts build in parallel).  See cov.cuh for the kernels, cov.cu for the C entry point.
//   M <= 8 : k_cov          -- 1-D split of the lower triangle over P compile-time parts (P = 1 for the bench shape)
//   M >= 9 : k_cov_blocked  -- BT x BT blocks with runtime block coordinates (one code path), source chunks <= 2
#include <stdlib.h>

#include "cov.cuh"
#include "cov_sweep.cuh"
#include "relayout_cov.cuh"

#ifndef OIVA_COV_M
#error "compile with -DOIVA_COV_M=<1..16>"
#endif

namespace oiva {

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

// frame splits per group that minimise the busiest team's work: ceil(units / teams) units of
// ceil(nchunks / split) chunks each, plus ~2 chunk-times of ring fill and write-out per unit
static int choose_split(long long G, int nchunks, long long n_teams, long long max_split) {
    long long hi = (16 * n_teams + G - 1) / G;  // (up to ~16 units per team: finer splits balance better)
    if (hi > nchunks) hi = nchunks;
    if (hi > max_split) hi = max_split < 1 ? 1 : max_split;
    long long best_cost = -1;
    int best = 1;
    for (long long ns = 1; ns <= hi; ++ns) {
        const long long per_team = (G * ns + n_teams - 1) / n_teams;
        const long long cost = per_team * ((nchunks + ns - 1) / ns + 2);
        if (best_cost < 0 || cost < best_cost) {
            best_cost = cost;
            best = (int)ns;
        }
    }
    return best;
}

// Every part of the 1-D entry split is its own unrolled code path: beyond 4 parts the kernel starves on instruction
// fetch (measured at M = K = 8, 256 mixtures: 8 sources per pass = 8 parts 17.6 ms; two passes of 4 sources = 4 parts
// each are 3x faster although X is read twice).  Larger splits are not instantiated.
constexpr int COV_MAX_PARTS = 4;
constexpr bool COV_USE_BLOCKS = OIVA_COV_M >= 9;

// shared launch logic: ring sizing, persistent grid, frame splitting for few groups
template <typename Kern, typename Launch>
static int launch_common(Kern kern, CovParams p, cudaStream_t st, int P, int teams_max, size_t stage_bytes, int TC,
                         int M, int KC, OivaPerDeviceOnce& attr_done, int* nsplit_out, Launch do_launch) {
    int dev = 0, sms = 148;
    OIVA_CUDA_CHECK(cudaGetDevice(&dev));
    OIVA_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int teams = env_int("OIVA_COV_TEAMS", teams_max);
    if (teams > teams_max) teams = teams_max;
    if (teams < 1) teams = 1;
    // measured on B200 (cfg4 shape): 2 stages x 8 single-warp teams per SM reaches 0.92 of the copy bandwidth,
    // 3-4 stages x 5 teams 0.70: more resident warps beat a deeper ring (profiles/r01_notes.md)
    const size_t budget = 200 * 1024;
    // ring depth: 2 stages for the 8 single-warp teams of the bench shape (their 16 stages fill the budget); shapes with
    // fewer, multi-warp teams (K >= 3 at M >= 6, M = 8, ...) leave shared memory unused at 2 stages -- they get as many
    // stages as fit, up to 4 (more chunks in flight per team: these teams are latency-bound, not bandwidth-bound)
    int S = env_int("OIVA_COV_STAGES", 0);
    if (S <= 0) {
        S = teams_max == 1 ? 4 : 2;
        const size_t bar = 128 * ((2 * 4 * sizeof(uint64_t) + 127) / 128);
        while (S < 4 && (size_t)teams * (bar + (size_t)(S + 1) * stage_bytes) <= budget) ++S;
    }
    if (S < 2) S = 2;
    size_t team_smem = 0, smem = 0;
    for (;;) {
        team_smem = 128 * ((2 * S * sizeof(uint64_t) + 127) / 128) + (size_t)S * stage_bytes;
        if (teams * team_smem <= budget) break;
        if (S > 3) { --S; continue; }
        if (teams > 1) { --teams; continue; }
        if (S > 2) { --S; continue; }
        oiva_set_error("oiva_weighted_cov: stage of %zu bytes does not fit shared memory", stage_bytes);
        return OIVA_ERR_INVALID;
    }
    smem = teams * team_smem;
    p.stages = S;
    const int threads = teams * P * 32;
    if (!attr_done[dev]) {
        OIVA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
        attr_done[dev] = true;
    }
    int occ = 1;
    OIVA_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
    if (occ < 1) occ = 1;
    const long long max_ctas = (long long)sms * occ;
    // few groups: split the frames of a group over several teams; the partial sums go to per-split slots of the
    // caller's scratch (summed in a fixed order afterwards) or, without scratch, atomically into a zeroed Vg.
    // All source-chunk passes of one call use the split count chosen by the first pass (p.nsplit preset > 0).
    const int nchunks = (p.L.T + TC - 1) / TC;
    if (p.nsplit <= 0) {
        p.nsplit = 1;
        if (p.G < 2 * max_ctas * teams && nchunks > 1) {
            p.nsplit = choose_split(p.G, nchunks, max_ctas * teams, p.Vpart ? p.max_split : nchunks);
        }
        if (p.nsplit > 1 && !p.Vpart)
            OIVA_CUDA_CHECK(cudaMemsetAsync(p.Vg, 0, (size_t)p.G * p.K * oiva_tri(M) * OIVA_GROUP * sizeof(cplx), st));
    }
    if (nsplit_out) *nsplit_out = p.nsplit;
    const long long U = p.G * p.nsplit;
    long long grid = (U + teams - 1) / teams;
    if (grid > max_ctas) grid = max_ctas;
    if (grid < 1) grid = 1;
    do_launch(p, (unsigned)grid, threads, smem, teams, (int)team_smem);
    OIVA_LAUNCH_CHECK();
    (void)KC;
    return OIVA_OK;
}

template <typename ST, int KC, bool WBIN = false>
static int launch(CovParams p, cudaStream_t st, int* nsplit_out) {
    constexpr int M = OIVA_COV_M;
    typedef typename StoreC<ST>::type XC;
    constexpr int TC = cov_chunk_frames(M);
    constexpr size_t x_stage = (size_t)TC * M * OIVA_GROUP * sizeof(XC);
    constexpr size_t stage_bytes =
        ((x_stage + (size_t)KC * TC * (WBIN ? OIVA_GROUP : 1) * sizeof(double) + 127) / 128) * 128;
    if constexpr (COV_USE_BLOCKS) {
        if constexpr (WBIN) {
            oiva_set_error("cov_launch: per-bin weights are implemented for M <= 8 (M=%d)", M);
            return OIVA_ERR_UNSUPPORTED;
        } else if constexpr (KC > 2) {
            oiva_set_error("cov_launch: blocked kernel (M=%d) takes source chunks of 1 or 2, not %d", M, KC);
            return OIVA_ERR_INVALID;
        } else {
            constexpr int P = cov_block_parts(M);
            auto kern = k_cov_blocked<ST, M, KC>;
            static OivaPerDeviceOnce attr_done;
            return launch_common(kern, p, st, P, 1, stage_bytes, TC, M, KC, attr_done, nsplit_out,
                                 [&](const CovParams& q, unsigned grid, int threads, size_t smem, int, int team_smem) {
                                     kern<<<grid, threads, smem, st>>>(q, team_smem);
                                 });
        }
    } else {
        constexpr int P = cov_parts(M, KC);
        // K <= M, so chunks larger than the smallest instantiated value >= M are never requested
        constexpr int kc_cap = M <= 4 ? M : (M <= 6 ? 6 : 8);
        if constexpr (P > COV_MAX_PARTS || KC > kc_cap) {
            oiva_set_error("cov_launch: (M=%d, KC=%d) is not instantiated (%d parts)", M, KC, P);
            return OIVA_ERR_INVALID;
        } else {
            auto kern = k_cov<ST, M, KC, P, WBIN>;
            static OivaPerDeviceOnce attr_done;
            constexpr int TCP = CovPart<ST, M, KC, P, 0, WBIN>::TC;  // (longer stages for four-warp teams)
            constexpr size_t x_stage_p = (size_t)TCP * M * OIVA_GROUP * sizeof(XC);
            constexpr size_t stage_bytes_p =
                ((x_stage_p + (size_t)KC * TCP * (WBIN ? OIVA_GROUP : 1) * sizeof(double) + 127) / 128) * 128;
            return launch_common(kern, p, st, P, cov_teams_per_cta(P), stage_bytes_p, TCP, M, KC, attr_done, nsplit_out,
                                 [&](const CovParams& q, unsigned grid, int threads, size_t smem, int teams,
                                     int team_smem) { kern<<<grid, threads, smem, st>>>(q, teams, team_smem); });
        }
    }
}

#define OIVA_CAT2(a, b) a##b
#define OIVA_CAT(a, b) OIVA_CAT2(a, b)

// largest usable source chunk for this M (register budget => number of parts)
int OIVA_CAT(cov_max_kc_m, OIVA_COV_M)() {
    constexpr int M = OIVA_COV_M;
    if (COV_USE_BLOCKS) return 2;
    int best = 1;
    if (cov_parts(M, 2) <= COV_MAX_PARTS) best = 2;
    if (cov_parts(M, 3) <= COV_MAX_PARTS) best = 3;
    if (cov_parts(M, 4) <= COV_MAX_PARTS) best = 4;
    if (cov_parts(M, 6) <= COV_MAX_PARTS) best = 6;
    if (cov_parts(M, 8) <= COV_MAX_PARTS) best = 8;
    return best;
}

int OIVA_CAT(cov_launch_m, OIVA_COV_M)(int dtype, int KC, const CovParams& p, cudaStream_t st, int* nsplit_out) {
#define OIVA_COV_CASE(ST_, KC_) \
    if (KC == KC_) return launch<ST_, KC_>(p, st, nsplit_out);
    if (dtype == OIVA_C64) {
        OIVA_COV_CASE(float, 1)
        OIVA_COV_CASE(float, 2)
        if (!COV_USE_BLOCKS) {
            OIVA_COV_CASE(float, 3)
            OIVA_COV_CASE(float, 4)
            OIVA_COV_CASE(float, 6)
            OIVA_COV_CASE(float, 8)
        }
    } else {
        OIVA_COV_CASE(double, 1)
        OIVA_COV_CASE(double, 2)
        if (!COV_USE_BLOCKS) {
            OIVA_COV_CASE(double, 3)
            OIVA_COV_CASE(double, 4)
            OIVA_COV_CASE(double, 6)
            OIVA_COV_CASE(double, 8)
        }
    }
    oiva_set_error("cov_launch: unsupported source chunk %d", KC);
    return OIVA_ERR_INVALID;
}

// per-bin weights (ILRMA): complex128 storage, M <= 8
int OIVA_CAT(cov_launch_wbin_m, OIVA_COV_M)(int KC, const CovParams& p, cudaStream_t st, int* nsplit_out) {
    if (COV_USE_BLOCKS) {
        oiva_set_error("cov_launch: per-bin weights are implemented for M <= 8");
        return OIVA_ERR_UNSUPPORTED;
    }
#define OIVA_COV_WCASE(KC_) \
    if (KC == KC_) return launch<double, KC_, true>(p, st, nsplit_out);
    OIVA_COV_WCASE(1)
    OIVA_COV_WCASE(2)
    OIVA_COV_WCASE(3)
    OIVA_COV_WCASE(4)
    OIVA_COV_WCASE(6)
    OIVA_COV_WCASE(8)
#undef OIVA_COV_WCASE
    oiva_set_error("cov_launch: unsupported source chunk %d", KC);
    return OIVA_ERR_INVALID;
}

// fused covariance + IP sweep (cov_sweep.cuh): single-warp teams, all K sources, no frame splits (many groups)
// (ring shape: measured at the bench shape with 2-frame stages x 4 and x 3, and 4-frame stages x 3 with fewer teams:
// 93.8 / 92.4 / 102.5 ms per step against 89.2 ms for the 4 frames x 2 stages x 8 teams of k_cov -- kept)
template <typename ST, int K>
static int launch_sweep(CovParams p, cudaStream_t st) {
    constexpr int M = OIVA_COV_M;
    if constexpr (!cov_sweep_supported(M, K)) {
        return OIVA_ERR_UNSUPPORTED;
    } else {
        typedef typename StoreC<ST>::type XC;
        constexpr int TC = cov_chunk_frames(M);
        constexpr size_t x_stage = (size_t)TC * M * OIVA_GROUP * sizeof(XC);
        constexpr size_t stage_bytes = ((x_stage + (size_t)K * TC * sizeof(double) + 127) / 128) * 128;
        auto kern = k_cov_sweep<ST, M, K>;
        static OivaPerDeviceOnce attr_done;
        p.nsplit = 1;
        return launch_common(kern, p, st, 1, cov_teams_per_cta(1), stage_bytes, TC, M, K, attr_done, nullptr,
                             [&](const CovParams& q, unsigned grid, int threads, size_t smem, int teams,
                                 int team_smem) { kern<<<grid, threads, smem, st>>>(q, teams, team_smem); });
    }
}

// OIVA_ERR_UNSUPPORTED (no error text) when (M, K) is not covered
int OIVA_CAT(cov_sweep_launch_m, OIVA_COV_M)(int dtype, int K, const CovParams& p, cudaStream_t st) {
#define OIVA_SWEEP_CASE(K_) \
    if (K == K_) return dtype == OIVA_C64 ? launch_sweep<float, K_>(p, st) : launch_sweep<double, K_>(p, st);
    OIVA_SWEEP_CASE(1)
    OIVA_SWEEP_CASE(2)
    OIVA_SWEEP_CASE(3)
#undef OIVA_SWEEP_CASE
    return OIVA_ERR_UNSUPPORTED;
}

// tiled kernel (cov.cuh: k_cov_tiled): M >= 9, chunks of 4 sources, clusters of two CTAs (the two halves of the tiles)
template <typename ST>
static int launch_tiled(CovParams p, cudaStream_t st, int* nsplit_out) {
    constexpr int M = OIVA_COV_M;
    if constexpr (M < 9) {
        oiva_set_error("cov_launch: the tiled kernel needs M >= 9");
        return OIVA_ERR_INVALID;
    } else {
        typedef typename StoreC<ST>::type XC;
        typedef CovTiling<M> TL;
        constexpr int TC = TL::TC, KC = TL::KC;
        constexpr size_t x_stage = (size_t)TC * M * OIVA_GROUP * sizeof(XC);
        constexpr size_t stage_bytes = ((x_stage + (size_t)KC * TC * sizeof(double) + 127) / 128) * 128;
        auto kern = k_cov_tiled<ST, M>;
        int S = env_int("OIVA_COV_TILED_STAGES", 3);
        if (S < 2) S = 2;
        const size_t budget = 200 * 1024;
        while (S > 2 && 128 * ((2 * S * sizeof(uint64_t) + 127) / 128) + (size_t)S * stage_bytes > budget) --S;
        const size_t smem = 128 * ((2 * S * sizeof(uint64_t) + 127) / 128) + (size_t)S * stage_bytes;
        p.stages = S;
        OIVA_SET_MAX_SMEM_ONCE(kern, budget);
        cudaLaunchConfig_t cfg = {};
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.blockDim = dim3(TL::WARPS * 32, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cfg.gridDim = dim3(2, 1, 1);
        int max_pairs = 0;
        OIVA_CUDA_CHECK(cudaOccupancyMaxActiveClusters(&max_pairs, kern, &cfg));
        if (max_pairs < 1) max_pairs = 1;
        const int nchunks = (p.L.T + TC - 1) / TC;
        if (p.nsplit <= 0) {
            p.nsplit = 1;
            if (p.G < 2ll * max_pairs && nchunks > 1 && p.Vpart && p.max_split > 1)
                p.nsplit = choose_split(p.G, nchunks, max_pairs, p.max_split);
        }
        if (p.nsplit > 1 && !p.Vpart) {
            oiva_set_error("cov_launch: the tiled kernel needs scratch for frame splits");
            return OIVA_ERR_INVALID;
        }
        if (nsplit_out) *nsplit_out = p.nsplit;
        long long pairs = p.G * p.nsplit;
        if (pairs > max_pairs) pairs = max_pairs;
        cfg.gridDim = dim3((unsigned)(2 * pairs), 1, 1);
        OIVA_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, p));
        return OIVA_OK;
    }
}

int OIVA_CAT(cov_launch_tiled_m, OIVA_COV_M)(int dtype, const CovParams& p, cudaStream_t st, int* nsplit_out) {
    return dtype == OIVA_C64 ? launch_tiled<float>(p, st, nsplit_out) : launch_tiled<double>(p, st, nsplit_out);
}

// relayout + input covariance in one pass (relayout_cov.cuh); M <= 8 only.  max_split: slots in p.Cpart (<= 1: none)
template <int M>
static int relayout_cov_launch_t(int dtype, RelayoutCovParams p, int max_split, cudaStream_t st, int* nsplit_out) {
    if constexpr (M > 8) {
        return OIVA_ERR_INVALID;
    } else {
        int dev = 0, sms = 148;
        OIVA_CUDA_CHECK(cudaGetDevice(&dev));
        OIVA_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        constexpr int TC = cov_chunk_frames(M);
        const size_t es = dtype == OIVA_C64 ? sizeof(float2) : sizeof(double2);
        const size_t stage_bytes = ((TC * OIVA_GROUP * M * es + 127) / 128) * 128;
        const int S = 2;
        int teams = 8;
        const size_t team_smem = 128 * ((2 * S * sizeof(uint64_t) + 127) / 128) + (size_t)S * stage_bytes;
        const size_t budget = 200 * 1024;
        while (teams > 1 && teams * team_smem > budget) --teams;
        const size_t smem = teams * team_smem;
        p.stages = S;
        const int threads = teams * 32;
        int occ = 1;
        if (dtype == OIVA_C64) {
            auto kern = k_relayout_cov<float, M>;
            OIVA_SET_MAX_SMEM_ONCE(kern, budget);
            OIVA_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
        } else {
            auto kern = k_relayout_cov<double, M>;
            OIVA_SET_MAX_SMEM_ONCE(kern, budget);
            OIVA_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
        }
        if (occ < 1) occ = 1;
        const long long max_ctas = (long long)sms * occ;
        const int nchunks = (p.L.T + TC - 1) / TC;
        p.nsplit = 1;
        if (p.G < 2 * max_ctas * teams && nchunks > 1 && p.Cpart && max_split > 1)
            p.nsplit = choose_split(p.G, nchunks, max_ctas * teams, max_split);
        if (nsplit_out) *nsplit_out = p.nsplit;
        const long long U = p.G * p.nsplit;
        long long grid = (U + teams - 1) / teams;
        if (grid > max_ctas) grid = max_ctas;
        if (grid < 1) grid = 1;
        if (dtype == OIVA_C64)
            k_relayout_cov<float, M><<<(unsigned)grid, threads, smem, st>>>(p, teams, (int)team_smem);
        else
            k_relayout_cov<double, M><<<(unsigned)grid, threads, smem, st>>>(p, teams, (int)team_smem);
        OIVA_LAUNCH_CHECK();
        return OIVA_OK;
    }
}

int OIVA_CAT(relayout_cov_launch_m, OIVA_COV_M)(int dtype, RelayoutCovParams p, int max_split, cudaStream_t st,
                                                int* nsplit_out) {
    return relayout_cov_launch_t<OIVA_COV_M>(dtype, p, max_split, st, nsplit_out);
}

}  // namespace oiva
