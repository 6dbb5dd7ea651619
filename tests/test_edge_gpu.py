"""Edge cases and full-size properties of the drop-in entry points on the GPU.

Edge shapes (single bin / frame counts below one staging chunk / one channel / the 16-channel maximum / bins not a
multiple of the 32-lane group) against the oracle; and, at BASELINE.json's full per-GPU size (512 mixtures of
(T=116, F=2049, M=6), K=2, 20 iterations), size-independent properties: bit-level determinism, equivariance under
a permutation of the batch, equivariance under a common rescaling of the input, and parity of sampled mixtures with
the oracle."""
import numpy as np
import pytest

from conftest import rel_err
from oracle import overiva_oracle as orc
from overiva_b200.synth import small_test_mixture, stft_domain_mixture

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if torch.cuda.is_available():
    import overiva_b200 as ob

TOL = 1e-10


# (the same list is pinned against the unmodified reference by oracle/validate_against_reference.py::EDGE_SHAPES)
@pytest.mark.parametrize("T,F,M,K,model", [
    (40, 1, 3, 2, "laplace"),      # a single frequency bin (31 padded lanes)
    (7, 33, 2, 2, "gauss"),        # fewer frames than one staging chunk; 33 bins = one full group + one bin
    (3, 5, 2, 1, "laplace"),       # T barely above M
    (50, 17, 1, 1, "laplace"),     # one channel: W is a scalar per bin
    (64, 9, 16, 16, "laplace"),    # the maximum channel count, determined
    (300, 9, 16, 4, "gauss"),      # BASELINE cfg5's (M, K) at a small (T, F)
    (90, 40, 9, 3, "laplace"),     # first channel count of the blocked covariance kernel
    (60, 64, 7, 7, "gauss"),       # row-owner sweep, determined
])
def test_edge_shapes(T, F, M, K, model):
    X = stft_domain_mixture(T * 1000 + F * 10 + M, T, F, M, K, n_interferers=min(3, max(0, M - K)), noise_db=-30.0)
    kw = dict(n_src=K, n_iter=5, model=model)
    Yo, Wo = orc.overiva(X, return_filters=True, **kw)
    Y, W = ob.overiva(X, return_filters=True, **kw)
    assert Y.shape == (T, F, K) and W.shape == (F, M, K)
    assert rel_err(Y, Yo) <= TOL and rel_err(W, Wo) <= TOL


def test_noncontiguous_and_readonly_inputs():
    X = small_test_mixture(31, 4, 2)
    big = np.zeros((X.shape[0], X.shape[1], 8), dtype=X.dtype)
    big[:, :, ::2] = X
    view = big[:, :, ::2]  # strided view, like X_all[:, :, :n_mics] in the drivers (overiva_oneshot.py:296)
    assert not view.flags["C_CONTIGUOUS"]
    ro = X.copy()
    ro.setflags(write=False)
    Yref = ob.overiva(X, n_src=2, n_iter=4)
    assert np.array_equal(ob.overiva(view, n_src=2, n_iter=4), Yref)  # same bits: the loop is deterministic
    assert np.array_equal(ob.overiva(ro, n_src=2, n_iter=4), Yref)


def test_input_validation():
    X = small_test_mixture(32, 3, 2)
    with pytest.raises(ValueError):
        ob.overiva(X[0], n_src=2)  # not 3-D
    with pytest.raises(TypeError):
        ob.overiva(X.real, n_src=2)  # not complex
    with pytest.raises(ValueError):
        ob.overiva(X, n_src=2, model="student")  # the one deliberate deviation: unknown models fail loudly
    with pytest.raises(ValueError):
        ob.overiva(np.zeros((10, 5, 17), dtype=np.complex128))  # more than 16 channels
    with pytest.raises(ValueError):
        ob.overiva(X, n_src=0)


def test_full_size_batch_properties():
    """BASELINE cfg4's per-GPU share.  No frame splitting / atomics at this size: results are bit-reproducible."""
    from overiva_b200.synth import stft_domain_batch_torch

    B, T, F, M, K = 512, 116, 2049, 6, 2
    dev = torch.device("cuda", torch.cuda.current_device())
    X = stft_domain_batch_torch(B, T, F, M, K, seed=2024, device=dev)
    kw = dict(n_src=K, n_iter=20, model="laplace")
    Y = ob.overiva_batch(X, **kw)
    assert Y.shape == (B, T, F, K) and bool(torch.isfinite(Y.real).all())
    # 1. determinism
    Y2 = ob.overiva_batch(X, **kw)
    assert torch.equal(Y, Y2)
    del Y2
    # 2. mixtures are independent: permuting the batch permutes the result, bit for bit
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(1)).to(dev)
    Yp = ob.overiva_batch(X[perm].contiguous(), **kw)
    assert torch.equal(Yp, Y[perm])
    del Yp
    # 3. projection back makes the output equivariant to a rescaling of the input (the laplace model's W is
    #    scale-invariant, Y = z W^H x scales with x): a power of two keeps every rounding identical
    Ys = ob.overiva_batch(X * 4.0, **kw)
    assert float((Ys - Y * 4.0).abs().max()) <= 1e-12 * float(Y.abs().max()) * 4.0
    del Ys
    # 4. sampled mixtures against the oracle
    for b in (0, 255, 511):
        Yo = orc.overiva(X[b].cpu().numpy(), **kw)
        assert rel_err(Y[b].cpu().numpy(), Yo) <= TOL
