// Instantiates the streaming demix kernels for ONE channel count (-DOIVA_M=<M>); see stream.cuh.
#include <type_traits>

#include "stream.cuh"

#ifndef OIVA_M
#error "compile with -DOIVA_M=<1..16>"
#endif

namespace oiva {

#define OIVA_CAT2(a, b) a##b
#define OIVA_CAT(a, b) OIVA_CAT2(a, b)

// smallest instantiated source chunk >= min(K, 8)
static int pick_kc(int K) {
    if (K <= 1) return 1;
    if (K == 2) return 2;
    if (K == 3) return 3;
    if (K == 4) return 4;
    return 8;
}

template <typename ST, int KC>
static int power_launch(StreamParams p, int n_batch, cudaStream_t st) {
    constexpr int M = OIVA_M;
    const int slots = p.L.frame_pitch() / 32;
    dim3 grid(p.NCH, oiva_div_up(slots, STREAM_WARPS), n_batch);
    for (int k0 = 0; k0 < p.K; k0 += KC) {
        p.k0 = k0;
        k_demix_power<ST, M, KC><<<grid, STREAM_WARPS * 32, 0, st>>>(p);
        OIVA_LAUNCH_CHECK();
    }
    return OIVA_OK;
}

template <typename ST, int KC>
static int output_launch(StreamParams p, int n_batch, cudaStream_t st) {
    constexpr int M = OIVA_M;
    const int slots = oiva_div_up(p.L.T, 32);
    dim3 grid(oiva_div_up(p.F, p.NBF), slots, n_batch);
    const size_t smem = (size_t)32 * (p.NBF * p.K + 1) * 2 * sizeof(ST);
    k_demix_output<ST, M, KC><<<grid, STREAM_WARPS * 32, smem, st>>>(p);
    OIVA_LAUNCH_CHECK();
    return OIVA_OK;
}

template <typename ST, int KC>
static int project_launch(StreamParams p, long long R, cudaStream_t st) {
    constexpr int M = OIVA_M;
    int fr = p.L.frame_pitch() > p.Lr.frame_pitch() ? p.L.frame_pitch() : p.Lr.frame_pitch();
    dim3 grid((unsigned)((R + STREAM_WARPS - 1) / STREAM_WARPS), fr / 32, 1);
    k_project_rows<ST, M, KC><<<grid, STREAM_WARPS * 32, 0, st>>>(p, R);
    OIVA_LAUNCH_CHECK();
    return OIVA_OK;
}

#define OIVA_KC_SWITCH(FN, ...)                                                              \
    switch (pick_kc(p.K)) {                                                                  \
        case 1: return dtype == OIVA_C64 ? FN<float, 1>(__VA_ARGS__) : FN<double, 1>(__VA_ARGS__); \
        case 2: return dtype == OIVA_C64 ? FN<float, 2>(__VA_ARGS__) : FN<double, 2>(__VA_ARGS__); \
        case 3: return dtype == OIVA_C64 ? FN<float, 3>(__VA_ARGS__) : FN<double, 3>(__VA_ARGS__); \
        case 4: return dtype == OIVA_C64 ? FN<float, 4>(__VA_ARGS__) : FN<double, 4>(__VA_ARGS__); \
        default: return dtype == OIVA_C64 ? FN<float, 8>(__VA_ARGS__) : FN<double, 8>(__VA_ARGS__); \
    }

int OIVA_CAT(power_launch_m, OIVA_M)(int dtype, const StreamParams& p, int n_batch, cudaStream_t st) {
    OIVA_KC_SWITCH(power_launch, p, n_batch, st)
}
int OIVA_CAT(output_launch_m, OIVA_M)(int dtype, const StreamParams& p, int n_batch, cudaStream_t st) {
    OIVA_KC_SWITCH(output_launch, p, n_batch, st)
}
int OIVA_CAT(project_launch_m, OIVA_M)(int dtype, const StreamParams& p, long long R, cudaStream_t st) {
    OIVA_KC_SWITCH(project_launch, p, R, st)
}

}  // namespace oiva
