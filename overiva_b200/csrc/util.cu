// Error plumbing, layout arithmetic, the (B,T,F,M) -> grouped-layout relayout kernel and the unpacking of
// grouped lower-triangle covariances into full row-major matrices.
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
extern "C" int oiva_version(void) { return 200; }

// ---- layout arithmetic (host only) --------------------------------------------------------------
static inline int elem_size(int dtype) { return dtype == OIVA_C64 ? 4 : 8; }

extern "C" int oiva_bin_groups(int n_freq) { return n_freq > 0 ? (n_freq + OIVA_GROUP - 1) / OIVA_GROUP : 0; }

extern "C" int oiva_frame_pitch(int n_frames) { return n_frames > 0 ? ((n_frames + 31) / 32) * 32 : 0; }

extern "C" size_t oiva_grouped_bytes(int n_batch, int n_frames, int n_freq, int n_chan, int dtype) {
    if (n_batch <= 0 || n_frames <= 0 || n_freq <= 0 || n_chan <= 0) return 0;
    return (size_t)n_batch * oiva_bin_groups(n_freq) * n_frames * n_chan * OIVA_GROUP * 2 * elem_size(dtype);
}

extern "C" size_t oiva_grouped_cov_bytes(int n_batch, int n_freq, int n_chan, int n_src) {
    if (n_batch <= 0 || n_freq <= 0 || n_chan <= 0 || n_src <= 0) return 0;
    return (size_t)n_batch * oiva_bin_groups(n_freq) * n_src * oiva_tri(n_chan) * OIVA_GROUP * 16;
}

// ---- relayout -----------------------------------------------------------------------------------
// grid (NG, ceil(T/FT), B), 128 threads = 4 warps; each warp transposes one frame of the group at a time:
// 32 bins x M channels (contiguous in X) -> M x 32 (contiguous in Xg) through a padded shared tile.
constexpr int RELAYOUT_FT = 32;  // frames per CTA
template <typename CT>
__global__ void __launch_bounds__(128) k_relayout(const CT* __restrict__ X, CT* __restrict__ Xg, GroupLayout L) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int M = L.M, T = L.T, F = L.F;
    const int pitch = M | 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    CT* tile = reinterpret_cast<CT*>(smem_raw) + (size_t)warp * 32 * pitch;
    const int g = blockIdx.x, b = blockIdx.z;
    const int f0 = g * OIVA_GROUP;
    const int nb = min(OIVA_GROUP, F - f0);
    const size_t gi = (size_t)b * L.NG + g;
    const int t_end = min(T, (int)(blockIdx.y + 1) * RELAYOUT_FT);
    for (int t = blockIdx.y * RELAYOUT_FT + warp; t < t_end; t += 4) {
        const CT* src = X + (((size_t)b * T + t) * F + f0) * M;
        for (int e = lane; e < 32 * M; e += 32) {
            const int l = e / M, c = e - l * M;
            CT v;
            v.x = 0;
            v.y = 0;
            if (l < nb) v = src[e];
            tile[l * pitch + c] = v;
        }
        __syncwarp();
        CT* dst = Xg + (gi * T + t) * (size_t)M * OIVA_GROUP;
        for (int c = 0; c < M; ++c) dst[c * OIVA_GROUP + lane] = tile[lane * pitch + c];
        __syncwarp();
    }
}

extern "C" int oiva_relayout(const void* X, void* Xg, int n_batch, int n_frames, int n_freq, int n_chan, int dtype,
                             void* stream) {
    OIVA_REQUIRE(X && Xg, "oiva_relayout: null pointer");
    OIVA_REQUIRE(n_batch > 0 && n_frames > 0 && n_freq > 0 && n_chan > 0 && n_chan <= OIVA_MAX_M,
                 "oiva_relayout: bad shape B=%d T=%d F=%d M=%d", n_batch, n_frames, n_freq, n_chan);
    OIVA_REQUIRE(n_batch <= 65535, "oiva_relayout: n_batch %d > 65535", n_batch);
    GroupLayout L = oiva_make_layout(n_frames, n_freq, n_chan);
    dim3 grid(L.NG, oiva_div_up(n_frames, RELAYOUT_FT), n_batch);
    OIVA_REQUIRE(grid.y <= 65535, "oiva_relayout: too many frames");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = (size_t)4 * 32 * (n_chan | 1) * 2 * elem_size(dtype);
    if (dtype == OIVA_C64)
        k_relayout<float2><<<grid, 128, smem, st>>>((const float2*)X, (float2*)Xg, L);
    else
        k_relayout<double2><<<grid, 128, smem, st>>>((const double2*)X, (double2*)Xg, L);
    OIVA_LAUNCH_CHECK();
    return OIVA_OK;
}

// ---- unpack grouped covariances -------------------------------------------------------------------
// Vg[gi][k][e][l] (lower triangle, bin-interleaved) -> V[row][k][i][j] full Hermitian, row = b*F + f.
__global__ void k_unpack_cov(const cplx* __restrict__ Vg, cplx* __restrict__ V, int F, int NG, int M, int K,
                             long long n) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const int j = (int)(idx % M), i = (int)((idx / M) % M);
    const int k = (int)((idx / ((long long)M * M)) % K);
    const long long row = idx / ((long long)M * M * K);
    const long long b = row / F;
    const int f = (int)(row - b * F);
    const size_t gi = (size_t)b * NG + f / OIVA_GROUP;
    const int l = f % OIVA_GROUP;
    const int NE = oiva_tri(M);
    const int hi = i >= j ? i : j, lo = i >= j ? j : i;
    cplx v = Vg[((gi * K + k) * NE + (hi * (hi + 1) / 2 + lo)) * OIVA_GROUP + l];
    if (i < j) v.y = -v.y;
    if (i == j) v.y = 0.0;
    V[idx] = v;
}

extern "C" int oiva_unpack_cov(const void* Vg, void* V, int n_batch, int n_freq, int n_chan, int n_src, void* stream) {
    OIVA_REQUIRE(Vg && V, "oiva_unpack_cov: null pointer");
    OIVA_REQUIRE(n_batch > 0 && n_freq > 0 && n_chan >= 1 && n_chan <= OIVA_MAX_M && n_src >= 1,
                 "oiva_unpack_cov: bad shape");
    const long long n = (long long)n_batch * n_freq * n_src * n_chan * n_chan;
    k_unpack_cov<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const cplx*)Vg, (cplx*)V, n_freq, oiva_bin_groups(n_freq), n_chan, n_src, n);
    OIVA_LAUNCH_CHECK();
    return OIVA_OK;
}

// ---- per-bin matrices: row-major (R, n) <-> grouped [gi][n][32] -------------------------------------
// The loop keeps the demixing matrices in the grouped form (lane <-> bin: a warp access to one matrix element of
// 32 consecutive bins is one 512-byte segment instead of 32 scattered 16-byte words).
__global__ void k_regroup_rows(const cplx* __restrict__ src, cplx* __restrict__ dst, int F, int NG, int n, long long G,
                               int to_grouped) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // over (gi, e, l)
    if (idx >= G * n * OIVA_GROUP) return;
    const int l = (int)(idx % OIVA_GROUP);
    const int e = (int)((idx / OIVA_GROUP) % n);
    const long long gi = idx / ((long long)OIVA_GROUP * n);
    const long long b = gi / NG;
    const int f = (int)(gi - b * NG) * OIVA_GROUP + l;
    if (to_grouped) {
        cplx v = make_double2(0.0, 0.0);
        if (f < F) v = src[((size_t)b * F + f) * n + e];
        dst[idx] = v;
    } else if (f < F) {
        dst[((size_t)b * F + f) * n + e] = src[idx];
    }
}

static int regroup(const void* src, void* dst, int n_batch, int n_freq, int n, int to_grouped, void* stream) {
    OIVA_REQUIRE(src && dst, "oiva_(un)group_rows: null pointer");
    OIVA_REQUIRE(n_batch > 0 && n_freq > 0 && n > 0, "oiva_(un)group_rows: bad shape");
    const int NG = oiva_bin_groups(n_freq);
    const long long G = (long long)n_batch * NG;
    const long long total = G * n * OIVA_GROUP;
    k_regroup_rows<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const cplx*)src, (cplx*)dst, n_freq,
                                                                                       NG, n, G, to_grouped);
    OIVA_LAUNCH_CHECK();
    return OIVA_OK;
}

extern "C" int oiva_group_rows(const void* rows, void* grouped, int n_batch, int n_freq, int n_elems, void* stream) {
    return regroup(rows, grouped, n_batch, n_freq, n_elems, 1, stream);
}
extern "C" int oiva_ungroup_rows(const void* grouped, void* rows, int n_batch, int n_freq, int n_elems, void* stream) {
    return regroup(grouped, rows, n_batch, n_freq, n_elems, 0, stream);
}
