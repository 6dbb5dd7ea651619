// Fused weighted covariance + IP sweep: the epilogue of the ONE pass over X that Appendix A of SURVEY.md describes
// (reference: overiva.py:176-190 -- per source a zgemm over X, then zgesv / normalise / J refresh per bin).
//
// For the shapes whose K lower triangles fit one lane's registers (single-warp teams, cov.cuh) the covariances of a
// bin never have to leave the lane that accumulated them: when the last frame of a group has been consumed the lane
// runs the thread-per-bin sweep of solve_tpb.cuh on its register copy of V_1..V_K and writes only the new W_hat.
// Against the two-kernel version (k_cov + k_ip_update_tpb) this drops the write and re-read of Vg (K * NE * 16 bytes
// per bin and epoch: 1.43 GB of the sweep's 1.97 GB at the bench shape) and one launch per epoch, and the sweep of
// one team overlaps the streaming of the other seven (the kernel stays HBM-bound).  The arithmetic is the same
// code in the same order, so W_hat is bit-identical to the two-kernel path.
#pragma once
#include "solve_tpb.cuh"  // (includes cov.cuh)

namespace oiva {

// single-warp teams (all K sources of the lower triangle in one lane's registers) and a thread-per-bin sweep
__host__ __device__ constexpr bool cov_sweep_supported(int M, int K) {
    return M >= 1 && M <= 8 && K >= 1 && K <= M && K <= 3 && cov_parts(M, K) == 1;
}

template <int M, int K>
struct CovSweepEpilogue {
    static constexpr int NE = oiva_tri(M);
    __device__ static __forceinline__ void run(cplx (&acc)[NE][K], const CovParams& p, int gi, int lane) {
        const int b = gi / p.L.NG;
        const int f = (gi - b * p.L.NG) * OIVA_GROUP + lane;
        if (f < p.L.F) {  // (padded lanes of a mixture's last group hold zeros)
            const WLane Wm = {p.Wg + (size_t)gi * M * M * OIVA_GROUP + lane};
            const cplx* Cl = p.Cg + (size_t)gi * NE * OIVA_GROUP + lane;
            bool singular = false;
            if (p.wscale) ip_sweep_rescale<M, K>(Wm, p.wscale + (size_t)b * K);
            const double invT = p.invT;
            static_for<K>([&](auto sc) {
                constexpr int s = decltype(sc)::value;
                cplx Vs[NE];  // exactly what CovPart::finish would have stored for source s
                static_for<NE>([&](auto ec) {
                    constexpr int e = decltype(ec)::value;
                    constexpr bool diag = ent_row(e) == ent_col(e);
                    Vs[e] = cmake(acc[e][s].x * invT, diag ? 0.0 : acc[e][s].y * invT);
                });
                if constexpr (K < M) {
                    ip_source_reduced_tri<M, K>(Wm, Vs, s, singular);
                    background_tpb<M, K, true>(Wm, Cl, singular);
                } else {
                    ip_source_full_v<M>(Wm, HermFromRegs<M>{Vs}, s, singular);
                }
            });
            const bool bad = ip_sweep_nonfinite<M, K>(Wm);
            if (singular || bad)
                atomicOr(p.status + b, (singular ? OIVA_STATUS_SINGULAR : 0) | (bad ? OIVA_STATUS_NONFINITE : 0));
        }
        __syncwarp();
    }
};

// blockDim.x = teams_per_cta * 32; dynamic smem = teams_per_cta * team_smem_bytes; launched with p.nsplit == 1
template <typename ST, int M, int K>
__global__ void __launch_bounds__(cov_threads(1)) k_cov_sweep(const CovParams p, int teams_per_cta, int team_smem_bytes) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int team = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* team_smem = smem_raw + (size_t)team * team_smem_bytes;
    if (lane == 0) {
        uint64_t* full = reinterpret_cast<uint64_t*>(team_smem);
        uint64_t* empty = full + p.stages;
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_fence_init();
    }
    __syncthreads();
    cov_team_body<CovPart<ST, M, K, 1, 0>, ST, M, K, CovSweepEpilogue<M, K>>(
        p, team_smem, (long long)blockIdx.x * teams_per_cta + team, (long long)gridDim.x * teams_per_cta, lane, true, 0);
}

}  // namespace oiva
