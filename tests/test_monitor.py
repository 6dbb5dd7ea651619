"""Device-side SDR / SIR monitor (reference: convergence_callback in overiva_oneshot.py:263-284).

CPU: the Gram-matrix form of the metric equals ``metrics.bss_eval``.  GPU: the Gram kernel against numpy (strided
inputs, lengths around the chunk size), and the monitor used as ``callback=`` against the same evaluation done on
the host with the oracle's synthesis."""
import numpy as np
import pytest

from oracle import overiva_oracle as orc
from oracle import stft_oracle as so
from overiva_b200 import metrics, monitor
from overiva_b200.synth import convolutive_mixture

torch = pytest.importorskip("torch")


@pytest.mark.parametrize("K,J", [(1, 1), (2, 2), (3, 4)])
def test_gram_form_equals_bss_eval(K, J):
    rng = np.random.default_rng(K * 10 + J)
    refs = rng.standard_normal((K, 4000))
    ests = rng.standard_normal((J, K)) @ refs + 0.3 * rng.standard_normal((J, 4000))
    X = np.concatenate([refs, ests])
    sdr, sir, perm = monitor.bss_eval_from_gram(X @ X.T, K)
    sdr0, sir0, perm0 = metrics.bss_eval(refs, ests)
    assert np.array_equal(perm, perm0)
    assert np.allclose(sdr, sdr0, atol=1e-9) and np.allclose(sir, sir0, atol=1e-9)


def test_monitor_needs_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        monitor.ConvergenceMonitor(np.zeros((2, 100)), framesize=64)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 255, 8192, 8193, 50001])
def test_gram_kernel(n):
    rng = np.random.default_rng(n)
    a = rng.standard_normal((3, n))
    yb = rng.standard_normal((n + 7, 4))  # channel-last audio, read in place through strides
    ad, yd = torch.from_numpy(a).cuda(), torch.from_numpy(yb).cuda()
    G = monitor.gram(ad, yd[5 : 5 + n, :2].T).cpu().numpy()
    X = np.concatenate([a, yb[5 : 5 + n, :2].T])
    assert np.allclose(G, X @ X.T, rtol=1e-12, atol=1e-12 * n)
    assert np.array_equal(G, G.T)
    G1 = monitor.gram(ad).cpu().numpy()
    assert np.allclose(G1, a @ a.T, rtol=1e-12, atol=1e-12 * n)
    assert np.array_equal(monitor.gram(ad).cpu().numpy(), G1)  # deterministic


@pytest.mark.gpu
def test_monitor_as_callback_matches_host_evaluation():
    import overiva_b200 as ob
    from overiva_b200 import stft as gst

    L_, hop = 64, 32
    mix, images = convolutive_mixture(5, 4, 2, duration=0.5, fs=8000, n_interferers=3, rt60=0.02, env_shape=2.0,
                                      env_block=0.02)
    wa = so.hann(L_)
    ws = so.compute_synthesis_window(wa, hop)
    X = so.analysis(mix, L_, hop, win=wa, pad_front=L_ - hop)
    mon = monitor.ConvergenceMonitor(images, framesize=L_, delay=L_ - hop)
    Y = ob.overiva(torch.from_numpy(X).cuda(), n_src=2, n_iter=25, callback=mon)
    assert Y.is_cuda and len(mon.SDR) == len(mon.SIR) == 3  # epochs 0, 10, 20 (overiva.py:142)
    # host evaluation of the same three estimates
    seen = []
    orc.overiva(X, n_src=2, n_iter=25, callback=lambda Yc: seen.append(Yc.copy()))
    for i, Yc in enumerate(seen):
        y = so.synthesis(Yc, L_, hop, win=ws)
        y = y[:, np.argsort(np.std(y, axis=0))[::-1]]
        m = min(y.shape[0] - (L_ - hop), images.shape[1])
        sdr, sir, _ = metrics.bss_eval(images[:, :m, 0], y[L_ - hop : L_ - hop + m, :2].T)
        assert np.allclose(mon.SDR[i], sdr, atol=1e-6) and np.allclose(mon.SIR[i], sir, atol=1e-6)
    assert np.mean(mon.SIR[-1]) > np.mean(mon.SIR[0])


@pytest.mark.gpu
def test_monitor_with_512_tap_filters_matches_host_bss_eval_sources():
    """filter_length = 512: the metric the reference's callback computes (mir_eval.separation.bss_eval_sources,
    overiva_oneshot.py:263-284), cross-correlations on the device, Toeplitz systems on the host."""
    import overiva_b200 as ob

    L_, hop = 256, 128
    mix, images = convolutive_mixture(6, 3, 2, duration=2.0, fs=8000, n_interferers=2, rt60=0.05, env_shape=2.0,
                                      env_block=0.02)
    wa = so.hann(L_)
    ws = so.compute_synthesis_window(wa, hop)
    X = so.analysis(mix, L_, hop, win=wa, pad_front=L_ - hop)
    mon = monitor.ConvergenceMonitor(images, framesize=L_, delay=L_ - hop, filter_length=512)
    ob.overiva(torch.from_numpy(X).cuda(), n_src=2, n_iter=11, callback=mon)
    assert len(mon.SDR) == 2
    seen = []
    orc.overiva(X, n_src=2, n_iter=11, callback=lambda Yc: seen.append(Yc.copy()))
    for i, Yc in enumerate(seen):
        y = so.synthesis(Yc, L_, hop, win=ws)
        y = y[:, np.argsort(np.std(y, axis=0))[::-1]]
        m = min(y.shape[0] - (L_ - hop), images.shape[1])
        sdr, sir, _, _ = metrics.bss_eval_sources(images[:2, :m, 0], y[L_ - hop : L_ - hop + m, :2].T, flen=512)
        assert np.allclose(mon.SDR[i], sdr, atol=1e-4) and np.allclose(mon.SIR[i], sir, atol=1e-4), (mon.SDR[i], sdr)
