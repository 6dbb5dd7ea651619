// Shared device helpers for the OverIVA sm_100a kernels: complex fp64 arithmetic on double2,
// mbarrier + 1-D bulk-TMA (cp.async.bulk) wrappers, warp utilities, the planar row layout.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/overiva_b200.h"

#define OIVA_MAX_M 16
#define OIVA_WARP 32

// ----------------------------------------------------------------------------------------------
// Grouped layout ("Xg") -- how the STFT lives in HBM for the whole loop: lane <-> frequency bin.
//
// The caller's X is (B, T, F, M) interleaved complex (the reference's (T, F, M), overiva.py:80); the
// reference re-lays it out as (F, T, M) (overiva.py:132).  Here the bins of every mixture are cut into
// NG = ceil(F/32) groups of 32 consecutive bins (the last group zero-padded) and one relayout kernel writes
//      Xg[gi = b*NG + g][t][c][l]      complex<ST>,  l = bin % 32
// Every per-bin quantity of the algorithm (covariances, demixing vectors) is then private to ONE LANE of a
// warp, so the streaming kernels need no cross-lane reduction for the covariance, a warp load of one
// (t, c) is 32 consecutive complex numbers (512 B in fp64: conflict-free LDS.128 / fully coalesced LDG.128),
// ANY range of frames of a group is one contiguous block (= one cp.async.bulk / TMA transaction, no tile
// structure in memory), and the (T, F, K)-ordered output is written as 32*K consecutive complex numbers
// per warp.  Weighted covariances use the same grouping: Vg[gi][k][e][l] for the lower-triangle entries
// e = i(i+1)/2 + j (i >= j).
// ----------------------------------------------------------------------------------------------
#define OIVA_GROUP 32
struct GroupLayout {
    int T;   // frames
    int F;   // bins per mixture
    int NG;  // groups per mixture = ceil(F / 32)
    int M;   // channels
    __host__ __device__ size_t frame_elems() const { return (size_t)M * OIVA_GROUP; }        // complex per frame
    __host__ __device__ size_t group_elems() const { return (size_t)T * M * OIVA_GROUP; }    // complex per group
    __host__ __device__ int frame_pitch() const { return ((T + 31) / 32) * 32; }             // phi / r2 rows
};
static inline GroupLayout oiva_make_layout(int n_frames, int n_freq, int n_chan) {
    GroupLayout L;
    L.T = n_frames;
    L.F = n_freq;
    L.NG = (n_freq + OIVA_GROUP - 1) / OIVA_GROUP;
    L.M = n_chan;
    return L;
}
__host__ __device__ constexpr int oiva_tri(int M) { return M * (M + 1) / 2; }

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
// 1 / x for the pivots of the sweeps: MUFU.RCP64H (23 bits) and two Newton steps, ~70 cycles of dependent latency where
// the IEEE division is ~200 (reciprocal seed, four refinements, a range check with a slow path); within 1 ulp for normal
// x.  x = 0 or Inf gives NaN instead of Inf / 0: both only occur behind a pivot test that has already flagged the bin.
__device__ __forceinline__ double rcp_fast(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}
__device__ __forceinline__ cplx crecip_fast(cplx a) {
    const double s = rcp_fast(fma(a.x, a.x, a.y * a.y));
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
// 1 / sqrt(d), principal branch, for the normalisation w^H V w of the sweeps (overiva.py:185-186): V is positive definite,
// so Re d > 0 and Im d is rounding noise.  sqrt(d) = (re, Im d / (2 re)) with re^2 = (|d| + Re d) / 2, and
// 1 / sqrt(d) = conj(sqrt(d)) / |d|: two rsqrt and a few multiplications (~220 cycles of latency) where hypot, two square
// roots and a division are ~700.  Anything else (Re d <= 0, NaN, squares out of range) takes the general formula.
__device__ __forceinline__ cplx crsqrt_pos(cplx d) {
    const double n2 = fma(d.x, d.x, d.y * d.y);
    if (!(d.x > 0.0) || !(n2 < 1e300) || !(n2 > 1e-300)) return crecip(csqrt_(d));
    const double rinv = rsqrt(n2);     // 1 / |d|
    const double t = 0.5 * fma(n2, rinv, d.x);  // re^2
    const double tinv = rsqrt(t);      // 1 / re
    const double re = t * tinv, im = 0.5 * d.y * tinv;
    return make_double2(re * rinv, -im * rinv);
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

// storage complex type for X / Y: double2 (c128) or float2 (c64); arithmetic is always fp64
template <typename ST> struct StoreC;
template <> struct StoreC<double> { typedef double2 type; };
template <> struct StoreC<float> { typedef float2 type; };
__device__ __forceinline__ cplx widen(double2 v) { return v; }
__device__ __forceinline__ cplx widen(float2 v) { return make_double2((double)v.x, (double)v.y); }
__device__ __forceinline__ void narrow(double2& o, cplx v) { o = v; }
__device__ __forceinline__ void narrow(float2& o, cplx v) { o = make_float2((float)v.x, (float)v.y); }
__device__ __forceinline__ cplx ldg_x(const double2* p) { return __ldg(p); }
__device__ __forceinline__ cplx ldg_x(const float2* p) { return widen(__ldg(p)); }

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

// non-blocking probe of a barrier phase
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

// ---- thread-block clusters: the same shared-memory offset in another CTA of the cluster, remote barrier arrival,
// cluster-wide barrier, multicast bulk copies (one L2 read lands in the shared memory of several CTAs) --------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_map(uint32_t smem_addr, uint32_t cta_rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(cta_rank));
    return r;
}
// (default semantics, as CUTLASS' ClusterBarrier::arrive: an explicit .release.cluster costs a MEMBAR per arrival --
// ncu showed 0.7 "membar" stalls per issued instruction in k_cov_tiled -- and is not needed to protect shared-memory
// READS that have already been consumed by arithmetic)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// probe of a barrier that also collects arrivals from other CTAs of the cluster
__device__ __forceinline__ bool mbar_test_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// one contiguous global block -> the SAME shared-memory offset of every CTA in cta_mask; each destination CTA's
// barrier (same offset) receives complete_tx for the bytes
__device__ __forceinline__ void tma_load_1d_multicast(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                                      uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::
            "r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask)
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
            oiva_set_error("%s:%d: %s -> CUDA error %d (%s)", __FILE__, __LINE__, #expr, (int)_e,  \
                           cudaGetErrorString(_e));                                             \
            return OIVA_ERR_CUDA;                                                               \
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

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a PER-DEVICE property of a kernel: launchers remember it per
// (kernel instantiation, device), not per process (one static instance per instantiation).
struct OivaPerDeviceOnce {
    bool done[64] = {};
    bool& operator[](int dev) { return done[(unsigned)dev & 63u]; }
};
#define OIVA_SET_MAX_SMEM_ONCE(kern, bytes)                                                                        \
    do {                                                                                                           \
        static OivaPerDeviceOnce once_;                                                                            \
        int dev_ = 0;                                                                                              \
        OIVA_CUDA_CHECK(cudaGetDevice(&dev_));                                                                     \
        if (!once_[dev_]) {                                                                                        \
            OIVA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
            once_[dev_] = true;                                                                                    \
        }                                                                                                          \
    } while (0)

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
