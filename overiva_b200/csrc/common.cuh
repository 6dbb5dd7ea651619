// Shared device helpers for the OverIVA sm_100a kernels: complex fp64 arithmetic on double2,
// mbarrier + 1-D bulk-TMA (cp.async.bulk) wrappers, warp utilities, the planar row layout.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/overiva_b200.h"

#define OIVA_MAX_M 16
#define OIVA_WARP 32

// ----------------------------------------------------------------------------------------------
// Planar row layout ("Xp") -- how the STFT lives in HBM for the whole loop.
//
// The caller's X is (B, T, F, M) interleaved complex (the reference's (T, F, M), overiva.py:80).
// The reference itself re-lays it out as (F, T, M) (overiva.py:132); here one relayout kernel writes,
// for every row = (mixture b, bin f), nT consecutive tiles
//      tile i : [plane = 2*c + ri][pitch_i frames]           (ST = double or float)
// with pitch_i = TT (a multiple of 32) for all but the last tile and TL = T - (nT-1)*TT for the last
// one (ragged: no padding in HBM, so tiling costs no bytes; for float storage TL is bumped by one zero
// frame when needed to keep every tile a multiple of 16 bytes).  For each channel c a tile holds a
// plane of real parts followed by a plane of imaginary parts.  Consecutive lanes of a warp read
// consecutive frames of one plane: every 32-lane load is one contiguous 256 B (fp64) segment,
// conflict-free in shared memory and fully coalesced in global memory, and a tile is one contiguous
// block, i.e. exactly one cp.async.bulk (TMA) transaction.
// ----------------------------------------------------------------------------------------------
struct RowLayout {
    int T;    // frames
    int TT;   // frames per full tile (multiple of 32)
    int nT;   // tiles per row
    int TL;   // allocated frames of the last tile (>= valid frames of the last tile)
    int M;    // channels
    __host__ __device__ int pitch(int tile) const { return tile + 1 < nT ? TT : TL; }
    __host__ __device__ int valid(int tile) const {
        int v = T - tile * TT;
        return v < TT ? v : TT;
    }
    __host__ __device__ size_t tile_off(int tile) const { return (size_t)tile * 2 * M * TT; }
    __host__ __device__ size_t row_elems() const { return (size_t)2 * M * ((size_t)(nT - 1) * TT + TL); }
    __host__ __device__ int frame_pitch() const { return nT * TT; }  // pitch of phi / r2 rows
};
RowLayout oiva_make_layout(int n_frames, int n_chan, int dtype);

// ----------------------------------------------------------------------------------------------
// complex helpers (double2: x = re, y = im)
// ----------------------------------------------------------------------------------------------
typedef double2 cplx;

__device__ __forceinline__ cplx cmake(double r, double i) { return make_double2(r, i); }
__device__ __forceinline__ cplx cconj(cplx a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx cscale(cplx a, double s) { return make_double2(a.x * s, a.y * s); }
// a * b
__device__ __forceinline__ cplx cmul(cplx a, cplx b) {
    return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
// conj(a) * b
__device__ __forceinline__ cplx cmulc(cplx a, cplx b) {
    return make_double2(fma(a.x, b.x, a.y * b.y), fma(a.x, b.y, -a.y * b.x));
}
// acc += a * b
__device__ __forceinline__ void cfma(cplx& acc, cplx a, cplx b) {
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(a.y, b.x, acc.y);
}
// acc += conj(a) * b
__device__ __forceinline__ void cfmac(cplx& acc, cplx a, cplx b) {
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(-a.y, b.x, acc.y);
}
// acc -= a * b
__device__ __forceinline__ void cfms(cplx& acc, cplx a, cplx b) {
    acc.x = fma(-a.x, b.x, acc.x);
    acc.x = fma(a.y, b.y, acc.x);
    acc.y = fma(-a.x, b.y, acc.y);
    acc.y = fma(-a.y, b.x, acc.y);
}
// 1 / a
__device__ __forceinline__ cplx crecip(cplx a) {
    double s = 1.0 / fma(a.x, a.x, a.y * a.y);
    return make_double2(a.x * s, -a.y * s);
}
// principal square root
__device__ __forceinline__ cplx csqrt_(cplx a) {
    if (a.y == 0.0) {
        if (a.x >= 0.0) return make_double2(sqrt(a.x), 0.0);
        return make_double2(0.0, sqrt(-a.x));
    }
    double m = hypot(a.x, a.y);
    if (a.x >= 0.0) {
        double re = sqrt(0.5 * (m + a.x));
        return make_double2(re, a.y / (2.0 * re));
    }
    double im = copysign(sqrt(0.5 * (m - a.x)), a.y);
    return make_double2(a.y / (2.0 * im), im);
}

__device__ __forceinline__ cplx shfl_c(cplx v, int src) {
    return make_double2(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}
__device__ __forceinline__ cplx shfl_xor_c(cplx v, int mask) {
    return make_double2(__shfl_xor_sync(0xffffffffu, v.x, mask), __shfl_xor_sync(0xffffffffu, v.y, mask));
}

__device__ __forceinline__ double ld_nc(const double* p) { return __ldg(p); }
__device__ __forceinline__ double ld_nc(const float* p) { return (double)__ldg(p); }
__device__ __forceinline__ cplx ld_nc_c(const cplx* p) { return __ldg(p); }

// ----------------------------------------------------------------------------------------------
// mbarrier + bulk TMA (cp.async.bulk, 1-D global -> shared) -- sm_90+/sm_100a PTX
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    // make the initialised barriers visible to the async (TMA) proxy
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "OIVA_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra OIVA_DONE;\n"
        "bra OIVA_WAIT;\n"
        "OIVA_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// one contiguous global block -> shared, completion signalled on `bar` (bytes: multiple of 16,
// both addresses 16-byte aligned)
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ----------------------------------------------------------------------------------------------
// host-side error plumbing
// ----------------------------------------------------------------------------------------------
void oiva_set_error(const char* fmt, ...);

#define OIVA_CUDA_CHECK(expr)                                                                   \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            oiva_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return (int)_e;                                                                     \
        }                                                                                       \
    } while (0)

#define OIVA_REQUIRE(cond, ...)             \
    do {                                    \
        if (!(cond)) {                      \
            oiva_set_error(__VA_ARGS__);    \
            return OIVA_ERR_INVALID;        \
        }                                   \
    } while (0)

#define OIVA_LAUNCH_CHECK() OIVA_CUDA_CHECK(cudaGetLastError())

static inline int oiva_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// dispatch a runtime channel count 1..16 to a compile-time template argument
#define OIVA_DISPATCH_M(Mval, ...)                                             \
    switch (Mval) {                                                            \
        case 1: { constexpr int M_ = 1; __VA_ARGS__; } break;                  \
        case 2: { constexpr int M_ = 2; __VA_ARGS__; } break;                  \
        case 3: { constexpr int M_ = 3; __VA_ARGS__; } break;                  \
        case 4: { constexpr int M_ = 4; __VA_ARGS__; } break;                  \
        case 5: { constexpr int M_ = 5; __VA_ARGS__; } break;                  \
        case 6: { constexpr int M_ = 6; __VA_ARGS__; } break;                  \
        case 7: { constexpr int M_ = 7; __VA_ARGS__; } break;                  \
        case 8: { constexpr int M_ = 8; __VA_ARGS__; } break;                  \
        case 9: { constexpr int M_ = 9; __VA_ARGS__; } break;                  \
        case 10: { constexpr int M_ = 10; __VA_ARGS__; } break;                \
        case 11: { constexpr int M_ = 11; __VA_ARGS__; } break;                \
        case 12: { constexpr int M_ = 12; __VA_ARGS__; } break;                \
        case 13: { constexpr int M_ = 13; __VA_ARGS__; } break;                \
        case 14: { constexpr int M_ = 14; __VA_ARGS__; } break;                \
        case 15: { constexpr int M_ = 15; __VA_ARGS__; } break;                \
        case 16: { constexpr int M_ = 16; __VA_ARGS__; } break;                \
        default: oiva_set_error("n_chan=%d unsupported (1..16)", (int)(Mval)); \
            return OIVA_ERR_INVALID;                                           \
    }
