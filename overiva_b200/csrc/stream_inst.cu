// Instantiates the streaming demix kernels for ONE channel count (-DOIVA_M=<M>); see stream.cuh.
#include <stdlib.h>

#include "stream.cuh"

#ifndef OIVA_M
#error "compile with -DOIVA_M=<1..16>"
#endif

namespace oiva {

#define OIVA_CAT2(a, b) a##b
#define OIVA_CAT(a, b) OIVA_CAT2(a, b)

// sources per warp: the lane keeps 2*M*KC doubles of filters in registers
// (M >= 13: two sources per warp -- 64 filter values -- halve the shared-memory reads of the staged kernels)
// (four sources per warp at M = 8 -- K = 8 as two chunks instead of three -- measured slower: 2.96 vs 2.70 ms per 256 mixtures)
constexpr int kc_cap() { return OIVA_M >= 13 ? 2 : ((24 / OIVA_M) > 4 ? 4 : (24 / OIVA_M)); }
static int pick_kc(int K) {
    // the fewest source chunks the register budget allows, then chunks of EQUAL size: the warps of a CTA share the
    // staged frames and wait for each other, so K = 6 runs as 3 + 3 sources, not 4 + 2
    const int cap = kc_cap() < 4 ? kc_cap() : 4;
    const int chunks = (K + cap - 1) / cap;
    return (K + chunks - 1) / chunks;
}

static int frame_splits(long long G, int units) {
    // enough CTAs for ~4 waves of 148 SMs x 4 resident CTAs, never more splits than units of work
    long long want = (4ll * 148 * 4 + G - 1) / G;
    if (want < 1) want = 1;
    if (want > units) want = units;
    if (want > 65535) want = 65535;
    return (int)want;
}

enum { KIND_POWER = 0, KIND_OUTPUT = 1, KIND_PROJECT = 2 };

template <typename ST, int KC>
static int launch(int kind, StreamParams p, long long G, cudaStream_t st) {
    constexpr int M = OIVA_M;
    if constexpr (KC > kc_cap()) {
        oiva_set_error("stream launch: KC=%d not instantiated for M=%d", KC, M);
        return OIVA_ERR_INVALID;
    } else {
        const int warps = (p.K + KC - 1) / KC;
        // many channels: one warp per group cannot hide the latency of its own loads (M = 16: 16 x LDG.128 per frame
        // and a 64-deep dependent chain) -- the staged kernels with 4 sub-block warps do better even for one source chunk
        const bool staged_single = M >= 9;
        const int units = kind == KIND_POWER ? p.L.frame_pitch() / POWER_FB : p.L.T;
        p.nsplit = frame_splits(G, units);
        dim3 grid((unsigned)G, p.nsplit, 1);
        static const bool no_staged = [] {
            const char* v = getenv("OIVA_POWER_NO_STAGE");
            return v && *v && *v != '0';
        }();
        if ((kind == KIND_POWER || kind == KIND_OUTPUT) && (warps > 1 || staged_single) && !no_staged) {
            // several warps per group: stage X once per CTA instead of once per warp; few source chunks are given
            // sub-blocks of the 8 frames so that a CTA still has up to 8 warps
            typedef typename StoreC<ST>::type XC;
            const size_t smem = 128 + 2 * (size_t)POWER_FB * M * OIVA_GROUP * sizeof(XC);
            p.nsplit = frame_splits(G, p.L.frame_pitch() / POWER_FB);
            grid = dim3((unsigned)G, p.nsplit, 1);
            const int nsub = warps >= 8 ? 1 : (warps >= 4 ? 2 : 4);
#define OIVA_STAGED(FBW_, OUT_)                                      \
    {                                                                \
        auto kern = k_demix_staged<ST, M, KC, FBW_, OUT_>;           \
        OIVA_SET_MAX_SMEM_ONCE(kern, smem);                          \
        kern<<<grid, 32 * warps * (POWER_FB / FBW_), smem, st>>>(p); \
    }
            if (kind == KIND_POWER) {
                if (nsub == 1) OIVA_STAGED(8, false) else if (nsub == 2) OIVA_STAGED(4, false) else OIVA_STAGED(2, false)
            } else {
                if (nsub == 1) OIVA_STAGED(8, true) else if (nsub == 2) OIVA_STAGED(4, true) else OIVA_STAGED(2, true)
            }
#undef OIVA_STAGED
        } else if (kind == KIND_POWER)
            k_demix_power<ST, M, KC><<<grid, 32 * warps, 0, st>>>(p);
        else if (kind == KIND_OUTPUT)
            k_demix_output<ST, M, KC><<<grid, 32 * warps, 0, st>>>(p);
        else
            k_project_rows<ST, M, KC><<<grid, 32 * warps, 0, st>>>(p);
        OIVA_LAUNCH_CHECK();
        return OIVA_OK;
    }
}

int OIVA_CAT(stream_launch_m, OIVA_M)(int kind, int dtype, const StreamParams& p, long long G, cudaStream_t st) {
    switch (pick_kc(p.K)) {
        case 1: return dtype == OIVA_C64 ? launch<float, 1>(kind, p, G, st) : launch<double, 1>(kind, p, G, st);
        case 2: return dtype == OIVA_C64 ? launch<float, 2>(kind, p, G, st) : launch<double, 2>(kind, p, G, st);
        case 3: return dtype == OIVA_C64 ? launch<float, 3>(kind, p, G, st) : launch<double, 3>(kind, p, G, st);
        default: return dtype == OIVA_C64 ? launch<float, 4>(kind, p, G, st) : launch<double, 4>(kind, p, G, st);
    }
}

}  // namespace oiva
