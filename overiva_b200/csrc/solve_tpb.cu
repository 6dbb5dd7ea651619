// Instantiations + launcher of the thread-per-bin IP sweep (solve_tpb.cuh) for M <= 6, K <= M and M = 7, 8 with K <= 4.
#include <stdlib.h>

#include "solve_pair.cuh"
#include "solve_tpb.cuh"

namespace oiva {

template <int M, int K>
static int launch_tpb(cplx* What, const cplx* Vg, const cplx* Cg, const double* wscale, int* status, int F, int NG,
                      long long G, cudaStream_t st) {
    const unsigned grid = (unsigned)((G + TPB_WARPS - 1) / TPB_WARPS);
    k_ip_update_tpb<M, K><<<grid, TPB_WARPS * 32, 0, st>>>(What, Vg, Cg, wscale, status, F, NG, G);
    OIVA_LAUNCH_CHECK();
    return OIVA_OK;
}

// returns OIVA_ERR_INVALID (without setting an error) when (M, K) is not covered
int ip_update_tpb(int M, int K, cplx* What, const cplx* Vg, const cplx* Cg, const double* wscale, int* status, int F,
                  int NG, long long G, cudaStream_t st) {
#define OIVA_TPB(M_, K_) \
    if (M == M_ && K == K_) return launch_tpb<M_, K_>(What, Vg, Cg, wscale, status, F, NG, G, st);
    OIVA_TPB(1, 1)
    OIVA_TPB(2, 1) OIVA_TPB(2, 2)
    OIVA_TPB(3, 1) OIVA_TPB(3, 2) OIVA_TPB(3, 3)
    OIVA_TPB(4, 1) OIVA_TPB(4, 2) OIVA_TPB(4, 3) OIVA_TPB(4, 4)
    OIVA_TPB(5, 1) OIVA_TPB(5, 2) OIVA_TPB(5, 3) OIVA_TPB(5, 4) OIVA_TPB(5, 5)
    OIVA_TPB(6, 1) OIVA_TPB(6, 2) OIVA_TPB(6, 3) OIVA_TPB(6, 4) OIVA_TPB(6, 5) OIVA_TPB(6, 6)
    // 7 and 8 channels: the overdetermined sweep only (Cholesky of V in registers + a K x K solve, K <= 4); the
    // determined LU of an 8 x 8 complex system does not fit one thread's registers and stays with the row-owner kernel
    OIVA_TPB(7, 1) OIVA_TPB(7, 2) OIVA_TPB(7, 3) OIVA_TPB(7, 4)
    OIVA_TPB(8, 1) OIVA_TPB(8, 2) OIVA_TPB(8, 3) OIVA_TPB(8, 4)
#undef OIVA_TPB
    return OIVA_ERR_INVALID;
}

// the determined sweep of 7 / 8 channels with two lanes per bin (solve_pair.cuh); OIVA_ERR_INVALID for other shapes
int ip_update_pair(int M, int K, cplx* Wg, const cplx* Vg, const double* wscale, int* status, int F, int NG, long long G,
                   cudaStream_t st) {
    if (K != M || (M != 7 && M != 8)) return OIVA_ERR_INVALID;
    const unsigned grid = (unsigned)((2 * G + PAIR_WARPS - 1) / PAIR_WARPS);
    if (M == 7) k_ip_update_pair<7><<<grid, PAIR_WARPS * 32, 0, st>>>(Wg, Vg, wscale, status, F, NG, G);
    else k_ip_update_pair<8><<<grid, PAIR_WARPS * 32, 0, st>>>(Wg, Vg, wscale, status, F, NG, G);
    OIVA_LAUNCH_CHECK();
    return OIVA_OK;
}

template <int M, int K>
static int launch_init(cplx* Wg, const cplx* Cg, const cplx* W0, int* status, int F, int NG, long long G, cudaStream_t st) {
    const unsigned grid = (unsigned)((G + TPB_WARPS - 1) / TPB_WARPS);
    k_init_grouped<M, K><<<grid, TPB_WARPS * 32, 0, st>>>(Wg, Cg, W0, status, F, NG, G);
    OIVA_LAUNCH_CHECK();
    return OIVA_OK;
}

// the grouped initialisation for the same (M, K) set; OIVA_ERR_UNSUPPORTED (no error text) when not covered
int init_demix_tpb(int M, int K, cplx* Wg, const cplx* Cg, const cplx* W0, int* status, int F, int NG, long long G,
                   cudaStream_t st) {
#define OIVA_TPB(M_, K_) \
    if (M == M_ && K == K_) return launch_init<M_, K_>(Wg, Cg, W0, status, F, NG, G, st);
    OIVA_TPB(1, 1)
    OIVA_TPB(2, 1) OIVA_TPB(2, 2)
    OIVA_TPB(3, 1) OIVA_TPB(3, 2) OIVA_TPB(3, 3)
    OIVA_TPB(4, 1) OIVA_TPB(4, 2) OIVA_TPB(4, 3) OIVA_TPB(4, 4)
    OIVA_TPB(5, 1) OIVA_TPB(5, 2) OIVA_TPB(5, 3) OIVA_TPB(5, 4) OIVA_TPB(5, 5)
    OIVA_TPB(6, 1) OIVA_TPB(6, 2) OIVA_TPB(6, 3) OIVA_TPB(6, 4) OIVA_TPB(6, 5) OIVA_TPB(6, 6)
    OIVA_TPB(7, 1) OIVA_TPB(7, 2) OIVA_TPB(7, 3) OIVA_TPB(7, 4)
    OIVA_TPB(8, 1) OIVA_TPB(8, 2) OIVA_TPB(8, 3) OIVA_TPB(8, 4)
#undef OIVA_TPB
    return OIVA_ERR_UNSUPPORTED;
}

}  // namespace oiva
