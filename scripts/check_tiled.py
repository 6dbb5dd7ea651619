#!/usr/bin/env python
"""Quick parity check of the tiled many-channel covariance kernel (cov.cuh: k_cov_tiled) against the blocked kernel
and the numpy oracle; exits non-zero on mismatch.  Used before long GPU sessions (a hang here is cheap)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import gpu_util as G  # noqa: E402
from oracle import overiva_oracle as orc  # noqa: E402
from overiva_b200.synth import small_test_mixture  # noqa: E402

worst = 0.0
for M, K, n_samples, frame, dtype in [(16, 4, 2300, 64, np.complex128), (13, 3, 3000, 32, np.complex128),
                                      (9, 7, 2100, 64, np.complex128), (12, 8, 1000, 64, np.complex64),
                                      (16, 16, 1500, 64, np.complex128), (10, 5, 40000, 64, np.complex128)]:
    B = 2
    X = np.stack([small_test_mixture(7 + b, M, 2, n_samples=n_samples, frame=frame, hop=frame // 2) for b in range(B)])
    X = X.astype(dtype)
    _, T, F, _ = X.shape
    rng = np.random.default_rng(5)
    phi = rng.gamma(1.0, 1.0, size=(B, K, T)) + 0.01
    Xg = G.grouped(X)
    V = G.weighted_cov(Xg, phi, B, T, F, M, K, G.code_of(dtype))
    X128 = X.astype(np.complex128)
    for b in range(B):
        Xf = np.ascontiguousarray(X128[b].swapaxes(0, 1))
        for s in range(K):
            want = orc.weighted_covariance(Xf, phi[b, s])
            err = np.linalg.norm(V[b, :, s] - want) / np.linalg.norm(want)
            worst = max(worst, err)
    print("M=%d K=%d T=%d F=%d %s: worst rel err so far %.2e" % (M, K, T, F, np.dtype(dtype).name, worst), flush=True)
assert worst < 1e-12, worst

# frame splits through the scratch slots (few groups, long mixture) + the whole loop against the oracle
import torch  # noqa: E402

import overiva_b200 as ob  # noqa: E402
from overiva_b200 import _lib as L  # noqa: E402

lib = L.load()
M, K = 16, 4
X = small_test_mixture(3, M, 2, n_samples=60000, frame=64, hop=32)  # T ~ 1870 frames, F = 33 -> 2 groups
T, F, _ = X.shape
rng = np.random.default_rng(6)
phi = rng.gamma(1.0, 1.0, size=(1, K, T)) + 0.01
Xg = G.grouped(X[None])
Tp = lib.oiva_frame_pitch(T)
ph = np.zeros((1, K, Tp))
ph[:, :, :T] = phi
phid = G.to_dev(ph)
nws = lib.oiva_weighted_cov_scratch_bytes(1, T, F, M, K)
ws = torch.empty(max(nws, 16), dtype=torch.uint8, device="cuda")
Vg = torch.full((lib.oiva_grouped_cov_bytes(1, F, M, K) // 8,), float("nan"), dtype=torch.float64, device="cuda")
L.check(lib.oiva_weighted_cov_ws(G.P(Xg), G.P(phid), G.P(Vg), G.P(ws), nws, 1, T, F, M, K, L.C128, G.stream()), "cov_ws")
V = torch.empty((1, F, K, M, M), dtype=torch.complex128, device="cuda")
L.check(lib.oiva_unpack_cov(G.P(Vg), G.P(V), 1, F, M, K, G.stream()), "unpack")
torch.cuda.synchronize()
V = V.cpu().numpy()
Xf = np.ascontiguousarray(X.swapaxes(0, 1))
for s in range(K):
    want = orc.weighted_covariance(Xf, phi[0, s])
    worst = max(worst, np.linalg.norm(V[0, :, s] - want) / np.linalg.norm(want))
print("split path (T=%d, scratch %d bytes): worst %.2e" % (T, nws, worst), flush=True)
assert worst < 1e-12, worst
Xs = small_test_mixture(11, M, 2, n_samples=9000, frame=128, hop=64)
Y, W = ob.overiva(Xs, n_src=K, n_iter=10, return_filters=True)
Yo, Wo = orc.overiva(Xs, n_src=K, n_iter=10, return_filters=True)
ey, ew = np.linalg.norm(Y - Yo) / np.linalg.norm(Yo), np.linalg.norm(W - Wo) / np.linalg.norm(Wo)
print("overiva M=16 K=4: rel err Y %.2e W %.2e" % (ey, ew), flush=True)
assert ey < 1e-10 and ew < 1e-10
print("TILED_OK")
