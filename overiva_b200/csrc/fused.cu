// Instantiations + launcher of the fused IP-sweep + next-epoch statistic kernel (fused.cuh), M <= 6, K <= 4.
#include "fused.cuh"

namespace oiva {

template <typename ST, int M, int K>
static int launch_fused(const FusedParams& p, cudaStream_t st) {
    k_solve_power<ST, M, K><<<(unsigned)((p.G + FUSED_WARPS - 1) / FUSED_WARPS), FUSED_WARPS * 32, 0, st>>>(p);
    OIVA_LAUNCH_CHECK();
    return OIVA_OK;
}

// OIVA_ERR_INVALID (no error text) when (M, K) is not covered: the caller falls back to the separate kernels
int solve_power_fused(int M, int K, int dtype, const FusedParams& p, cudaStream_t st) {
#define OIVA_F(M_, K_)                                                                  \
    if (M == M_ && K == K_)                                                             \
        return dtype == OIVA_C64 ? launch_fused<float, M_, K_>(p, st) : launch_fused<double, M_, K_>(p, st);
    OIVA_F(2, 1) OIVA_F(2, 2)
    OIVA_F(3, 1) OIVA_F(3, 2) OIVA_F(3, 3)
    OIVA_F(4, 1) OIVA_F(4, 2) OIVA_F(4, 3) OIVA_F(4, 4)
    OIVA_F(5, 1) OIVA_F(5, 2) OIVA_F(5, 3) OIVA_F(5, 4)
    OIVA_F(6, 1) OIVA_F(6, 2) OIVA_F(6, 3) OIVA_F(6, 4)
#undef OIVA_F
    return OIVA_ERR_INVALID;
}

}  // namespace oiva

using namespace oiva;

extern "C" int oiva_ip_update_power(void* Wg, const void* Vg, const void* Cg, const double* wscale, int* status,
                                    const void* Xg, double* r2part, int n_batch, int n_frames, int n_freq, int n_chan,
                                    int n_src, int dtype, void* stream) {
    OIVA_REQUIRE(Wg && Vg && Cg && status && Xg && r2part, "oiva_ip_update_power: null pointer");
    OIVA_REQUIRE(n_batch > 0 && n_frames > 0 && n_freq > 0 && n_src >= 1 && n_src <= n_chan,
                 "oiva_ip_update_power: bad shape");
    FusedParams p;
    p.Wg = (cplx*)Wg;
    p.Vg = (const cplx*)Vg;
    p.Cg = (const cplx*)Cg;
    p.wscale = wscale;
    p.status = status;
    p.Xg = Xg;
    p.r2part = r2part;
    p.L = oiva_make_layout(n_frames, n_freq, n_chan);
    p.G = (long long)n_batch * p.L.NG;
    int rc = solve_power_fused(n_chan, n_src, dtype, p, (cudaStream_t)stream);
    if (rc == OIVA_ERR_INVALID) oiva_set_error("oiva_ip_update_power: (M=%d, K=%d) not covered (M <= 6, K <= 4)", n_chan, n_src);
    return rc;
}

// 1 when oiva_ip_update_power covers (n_chan, n_src)
extern "C" int oiva_ip_update_power_supported(int n_chan, int n_src) {
    return n_chan >= 2 && n_chan <= 6 && n_src >= 1 && n_src <= n_chan && n_src <= 4;
}
