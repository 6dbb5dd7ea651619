// Error plumbing, layout arithmetic and the (B,T,F,M) -> planar-rows relayout kernel.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void oiva_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* oiva_last_error(void) { return g_err; }
extern "C" int oiva_version(void) { return 100; }

// ---- layout arithmetic --------------------------------------------------------------------------
static inline int elem_size(int dtype) { return dtype == OIVA_C64 ? 4 : 8; }

RowLayout oiva_make_layout(int n_frames, int n_chan, int dtype) {
    RowLayout L;
    L.T = n_frames;
    L.M = n_chan;
    L.TT = n_chan <= 8 ? 128 : 64;  // a full tile is <= 16 KB in fp64
    L.nT = (n_frames + L.TT - 1) / L.TT;
    if (L.nT < 1) L.nT = 1;
    int tl = n_frames - (L.nT - 1) * L.TT;
    // every tile must be a multiple of 16 bytes for the bulk copy (always true in fp64)
    while (((size_t)2 * n_chan * tl * elem_size(dtype)) % 16 != 0) ++tl;
    L.TL = tl;
    return L;
}

extern "C" int oiva_tile_frames(int n_frames, int n_chan, int dtype) {
    if (n_frames <= 0 || n_chan <= 0) return 0;
    return oiva_make_layout(n_frames, n_chan, dtype).TT;
}

extern "C" int oiva_frame_pitch(int n_frames, int n_chan, int dtype) {
    if (n_frames <= 0 || n_chan <= 0) return 0;
    return oiva_make_layout(n_frames, n_chan, dtype).frame_pitch();
}

extern "C" size_t oiva_planar_bytes(int n_batch, int n_frames, int n_freq, int n_chan, int dtype) {
    if (n_frames <= 0 || n_chan <= 0) return 0;
    RowLayout L = oiva_make_layout(n_frames, n_chan, dtype);
    return (size_t)n_batch * n_freq * L.row_elems() * elem_size(dtype);
}

extern "C" int oiva_power_chunks(int n_batch, int n_freq) {
    if (n_batch <= 0 || n_freq <= 0) return 0;
    long long bins = (long long)n_batch * n_freq;
    long long per = bins / (148 * 8);
    if (per < 8) per = 8;
    if (per > 64) per = 64;
    return (int)((n_freq + per - 1) / per);
}

// ---- relayout -----------------------------------------------------------------------------------
// grid (ceil(F/FB), nT*TT/32, B), 256 threads; smem tile [32][FB*2M + 1]
template <typename ST>
__global__ void __launch_bounds__(256) k_relayout(const ST* __restrict__ X, ST* __restrict__ Xp, RowLayout L, int F,
                                                  int FB) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ST* tile = reinterpret_cast<ST*>(smem_raw);
    const int M = L.M, T = L.T;
    const int f0 = blockIdx.x * FB;
    const int t0 = blockIdx.y * 32;
    const int b = blockIdx.z;
    const int nb = min(FB, F - f0);
    const int J = nb * 2 * M;
    const int pitch = FB * 2 * M + 1;
    const int tid = threadIdx.x;

    for (int i = tid; i < 32 * J; i += 256) {
        int tl = i / J, j = i - tl * J;
        int t = t0 + tl;
        ST v = (ST)0;
        if (t < T) v = X[(((size_t)b * T + t) * F + f0) * 2 * M + j];
        tile[tl * pitch + j] = v;
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    const int t = t0 + lane;
    const int ti = t / L.TT;  // t0 is a multiple of 32 and so is TT: the whole block lies in one tile
    if (ti >= L.nT) return;
    const int tloc = t - ti * L.TT;
    const int tp = L.pitch(ti);
    if (tloc >= tp) return;
    const size_t row_elems = L.row_elems();
    for (int j = warp; j < J; j += 8) {
        int fb = j / (2 * M);
        int plane = j - fb * 2 * M;  // 2*c + ri
        size_t row = (size_t)b * F + f0 + fb;
        Xp[row * row_elems + L.tile_off(ti) + (size_t)plane * tp + tloc] = tile[lane * pitch + j];
    }
}

extern "C" int oiva_relayout(const void* X, void* Xp, int n_batch, int n_frames, int n_freq, int n_chan, int dtype,
                             void* stream) {
    OIVA_REQUIRE(X && Xp, "oiva_relayout: null pointer");
    OIVA_REQUIRE(n_batch > 0 && n_frames > 0 && n_freq > 0 && n_chan > 0 && n_chan <= OIVA_MAX_M,
                 "oiva_relayout: bad shape B=%d T=%d F=%d M=%d", n_batch, n_frames, n_freq, n_chan);
    OIVA_REQUIRE(n_batch <= 65535, "oiva_relayout: n_batch %d > 65535", n_batch);
    RowLayout L = oiva_make_layout(n_frames, n_chan, dtype);
    int FB = 128 / (2 * n_chan);
    if (FB < 1) FB = 1;
    dim3 grid(oiva_div_up(n_freq, FB), L.nT * L.TT / 32, n_batch);
    OIVA_REQUIRE(grid.y <= 65535, "oiva_relayout: too many frames");
    cudaStream_t st = (cudaStream_t)stream;
    size_t smem = (size_t)32 * (FB * 2 * n_chan + 1) * elem_size(dtype);
    if (dtype == OIVA_C64)
        k_relayout<float><<<grid, 256, smem, st>>>((const float*)X, (float*)Xp, L, n_freq, FB);
    else
        k_relayout<double><<<grid, 256, smem, st>>>((const double*)X, (double*)Xp, L, n_freq, FB);
    OIVA_LAUNCH_CHECK();
    return OIVA_OK;
}
