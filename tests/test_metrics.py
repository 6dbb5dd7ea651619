"""SDR / SIR / SAR (overiva_b200.metrics): the one-tap decomposition used on both sides of the 0.1 dB parity test and by
the device monitor, and the restatement of mir_eval's bss_eval_sources (512-tap distortion filters) for final scoring."""
import numpy as np
import pytest

from overiva_b200 import metrics


def _sources(K, N, seed):
    rng = np.random.default_rng(seed)
    s = rng.laplace(size=(K, N))
    # give the sources some colour so that delayed copies are not orthogonal
    for k in range(K):
        s[k] = np.convolve(s[k], rng.standard_normal(8))[:N]
        s[k] /= np.std(s[k])
    return s, rng


def test_one_tap_filters_reduce_to_bss_eval():
    refs, rng = _sources(3, 3000, 0)
    ests = rng.standard_normal((3, 3)) @ refs + 0.2 * rng.standard_normal((3, 3000))
    sdr, sir, sar, perm = metrics.bss_eval_sources(refs, ests, flen=1)
    sdr0, sir0, perm0 = metrics.bss_eval(refs, ests)
    assert np.array_equal(perm, perm0)
    assert np.allclose(sdr, sdr0, atol=1e-9) and np.allclose(sir, sir0, atol=1e-9)


def test_filtered_target_is_not_a_distortion():
    """An estimate that is a (<= flen tap) filtered version of its source plus a filtered interferer has no artifacts:
    SAR is huge, SDR == SIR, and the one-tap metric wrongly counts the filtering as distortion."""
    refs, rng = _sources(2, 4000, 1)
    refs[:, -40:] = 0.0  # so that the 16-tap filtered signals are complete within N samples (no truncated tail)
    h = rng.standard_normal(16)
    g = 0.1 * rng.standard_normal(16)
    est0 = np.convolve(refs[0], h)[:4000] + np.convolve(refs[1], g)[:4000]
    est1 = np.convolve(refs[1], h[::-1])[:4000] + np.convolve(refs[0], g)[:4000]
    sdr, sir, sar, perm = metrics.bss_eval_sources(refs, np.stack([est0, est1]), flen=32)
    assert list(perm) == [0, 1]
    assert np.all(sar > 60) and np.allclose(sdr, sir, atol=1e-3)
    assert np.all(sir > 10) and np.all(sir < 40)
    sdr1, _, _ = metrics.bss_eval(refs, np.stack([est0, est1]))
    assert np.all(sdr1 < sdr - 3)


def test_artifacts_and_permutation():
    refs, rng = _sources(2, 3000, 2)
    noise = rng.standard_normal((2, 3000))
    ests = np.stack([refs[1] + 0.05 * refs[0] + 0.1 * noise[0], refs[0] + 0.05 * refs[1] + 0.1 * noise[1]])  # swapped
    sdr, sir, sar, perm = metrics.bss_eval_sources(refs, ests, flen=8)
    assert list(perm) == [1, 0]
    # SIR ~ 20 log10(1 / 0.05) = 26 dB (colouring changes it a little), SAR ~ 20 dB for noise at -20 dB
    assert np.all(np.abs(sir - 26) < 3) and np.all(np.abs(sar - 20) < 3) and np.all(sdr < sar)
    with pytest.raises(ValueError):
        metrics.bss_eval_sources(refs, ests[:1])


def test_delayed_gram_matches_direct_computation():
    refs, _ = _sources(2, 300, 3)
    flen = 5
    G = metrics._delayed_gram(refs, flen)
    pad = np.zeros((2, flen - 1))
    D = np.stack([np.concatenate([np.zeros(a), refs[k], np.zeros(flen - 1 - a)]) for k in range(2) for a in range(flen)])
    assert np.allclose(G, D @ D.T, atol=1e-9)
    assert pad.shape == (2, 4)


@pytest.mark.parametrize("flen,K", [(1, 2), (8, 2), (64, 3)])
def test_bss_eval_sources_from_cross_correlations(flen, K):
    """Every quantity of the filter-based metric is a function of cross-correlations at flen lags: the route the device
    monitor takes (oiva_xcorr + host solves) gives the same numbers as the signal-domain restatement."""
    refs, rng = _sources(K, 2500, 10 + flen)
    ests = rng.standard_normal((K, K)) @ refs + 0.3 * rng.standard_normal((K, 2500))
    ests[0] = np.convolve(ests[0], rng.standard_normal(4))[:2500]
    c, ee = metrics.xcorr_reference(refs, ests, flen)
    # the definition, directly
    for i, j, m in [(0, 1, 0), (K - 1, K, flen - 1), (1, 0, flen // 2)]:
        sj = np.concatenate([refs, ests])[j]
        assert np.isclose(c[i, j, m], np.dot(refs[i][: 2500 - m], sj[m:]), rtol=1e-9, atol=1e-9)
    got = metrics.bss_eval_sources_from_xcorr(c, ee, flen)
    want = metrics.bss_eval_sources(refs, ests, flen=flen)
    assert np.array_equal(got[3], want[3])
    for a, b in zip(got[:3], want[:3]):
        assert np.allclose(a, b, atol=1e-6), (a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("flen,N", [(1, 1000), (32, 5000), (512, 40000), (1000, 3000)])
def test_device_xcorr_and_filter_metric(flen, N):
    import torch

    from overiva_b200 import monitor

    K = 2
    refs, rng = _sources(K, N, 20 + flen)
    ests = rng.standard_normal((K, K)) @ refs + 0.2 * rng.standard_normal((K, N))
    rd, ed = torch.from_numpy(refs).cuda(), torch.from_numpy(ests.T.copy()).cuda().T  # (estimates with a sample-major stride)
    c = monitor.xcorr(rd, ed, flen).cpu().numpy()
    cw, ee = metrics.xcorr_reference(refs, ests, flen)
    assert c.shape == cw.shape and np.allclose(c, cw, rtol=1e-10, atol=1e-8 * np.abs(cw).max())
    assert np.array_equal(c, monitor.xcorr(rd, ed, flen).cpu().numpy())  # deterministic
    if flen <= 512:
        got = monitor.bss_eval_sources_device(rd, ed, flen)
        want = metrics.bss_eval_sources(refs, ests, flen=flen)
        assert np.array_equal(got[3], want[3])
        for a, b in zip(got[:3], want[:3]):
            assert np.allclose(a, b, atol=1e-5), (a, b)
