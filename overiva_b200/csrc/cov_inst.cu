// Instantiates the covariance kernels for ONE channel count (compiled once per M with -DOIVA_COV_M=<M>,
// so the 16 channel counts build in parallel).  See cov.cuh for the kernel, cov.cu for the C entry point.
#include <stdlib.h>

#include "cov.cuh"

#ifndef OIVA_COV_M
#error "compile with -DOIVA_COV_M=<1..16>"
#endif

namespace oiva {

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

template <typename ST, int KC, bool TMA>
static int launch(CovParams p, cudaStream_t st) {
    constexpr int M = OIVA_COV_M;
    constexpr int P = cov_parts(M, KC);
    static_assert(P <= 32, "team too large");
    auto kern = k_cov<ST, M, KC, P, TMA>;

    int dev = 0, sms = 148;
    OIVA_CUDA_CHECK(cudaGetDevice(&dev));
    OIVA_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));

    const RowLayout& L = p.L;
    p.xpitch = L.nT > 1 ? L.TT : L.TL;
    p.ppitch = (p.xpitch + 1) & ~1;
    const size_t x_stage = (size_t)2 * M * p.xpitch * sizeof(ST);
    const size_t stage_bytes = ((x_stage + (size_t)KC * p.ppitch * sizeof(double) + 127) / 128) * 128;
    int teams = env_int("OIVA_COV_TEAMS", cov_teams_per_cta(P));
    if (teams > cov_teams_per_cta(P)) teams = cov_teams_per_cta(P);
    if (teams < 1) teams = 1;
    int S = env_int("OIVA_COV_STAGES", 3);
    if (S < 2) S = 2;
    size_t team_smem = 0, smem = 0;
    const size_t budget = 200 * 1024;
    if (TMA) {
        for (;;) {
            team_smem = 128 * ((2 * S * sizeof(uint64_t) + 127) / 128) + (size_t)S * stage_bytes;
            if (teams * team_smem <= budget) break;
            if (S > 2) { --S; continue; }
            if (teams > 1) { --teams; continue; }
            oiva_set_error("oiva_weighted_cov: tile of %zu bytes does not fit shared memory", stage_bytes);
            return OIVA_ERR_INVALID;
        }
        smem = teams * team_smem;
    }
    p.stages = S;
    const int threads = teams * P * 32;

    static bool attr_done = false;  // per instantiation
    if (!attr_done) {
        OIVA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
        attr_done = true;
    }
    int occ = 1;
    OIVA_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
    if (occ < 1) occ = 1;
    long long max_ctas = (long long)sms * occ;
    // few rows and long rows: split the tiles of a row over several teams (atomic accumulation)
    p.nsplit = 1;
    if ((long long)p.R < max_ctas * teams && L.nT > 1) {
        long long want = (max_ctas * teams + p.R - 1) / p.R;
        p.nsplit = (int)(want < L.nT ? want : L.nT);
    }
    if (p.nsplit > 1 && p.k0 == 0)
        OIVA_CUDA_CHECK(cudaMemsetAsync(p.V, 0, (size_t)p.R * p.K * M * M * 2 * sizeof(double), st));
    const long long U = (long long)p.R * p.nsplit;
    long long grid = (U + teams - 1) / teams;
    if (grid > max_ctas) grid = max_ctas;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, threads, smem, st>>>(p, teams, (int)team_smem);
    OIVA_LAUNCH_CHECK();
    if (p.nsplit > 1) {
        const long long n = (long long)p.R * KC * M * M;
        k_cov_mirror<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p.V, p.R, p.K, p.k0, KC, M);
        OIVA_LAUNCH_CHECK();
    }
    return OIVA_OK;
}

#define OIVA_CAT2(a, b) a##b
#define OIVA_CAT(a, b) OIVA_CAT2(a, b)

int OIVA_CAT(cov_launch_m, OIVA_COV_M)(int dtype, int KC, int use_tma, const CovParams& p, cudaStream_t st) {
#define OIVA_COV_CASE(ST_, KC_)                                  \
    if (KC == KC_) {                                             \
        if (use_tma) return launch<ST_, KC_, true>(p, st);       \
        return launch<ST_, KC_, false>(p, st);                   \
    }
    if (dtype == OIVA_C64) {
        OIVA_COV_CASE(float, 1)
        OIVA_COV_CASE(float, 2)
        OIVA_COV_CASE(float, 4)
    } else {
        OIVA_COV_CASE(double, 1)
        OIVA_COV_CASE(double, 2)
        OIVA_COV_CASE(double, 4)
    }
    oiva_set_error("cov_launch: unsupported source chunk %d", KC);
    return OIVA_ERR_INVALID;
}

}  // namespace oiva
