"""Randomised parity on the GPU (tools/fuzz_parity.py): random voices over the reference GUI's ranges, random-walk tracks
with edge values patched in (parameters exactly at amplitude()'s clamps, closed velum, frication tap at the tube's ends,
pitch extremes), ragged lengths, both output rates, mono and stereo -- the three arithmetic modes against each other and
against the CPU oracle."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


@pytest.mark.parametrize("seed", [11, 12])
def test_random_voices_and_edge_tracks(seed):
    import fuzz_parity
    worst, failures = fuzz_parity.run(n=160, seed=seed, n_or=8)
    assert failures == 0
    assert worst["cs"][0] <= 1e-9, "conformance vs strict, utterance %d" % worst["cs"][1]          # north_star tolerance
    assert worst["pcm"][0] <= 1, "PCM, utterance %d" % worst["pcm"][1]
    assert worst["mx"][0] <= 1e-9
    assert worst["so"][0] <= 1e-9, "strict vs oracle, utterance %d" % worst["so"][1]
    assert worst["snr"][0] >= 60.0, "FP32 fast mode, utterance %d" % worst["snr"][1]


@pytest.mark.parametrize("precision", [0, 1, 2])
@pytest.mark.parametrize("seed", [21, 22, 23])
def test_random_pushes_of_random_voices_concatenate_bit_identically(precision, seed):
    """TRMStream with a random voice (either converter direction) and random push sizes: the concatenated pushes equal the
    one-shot result bit for bit in every arithmetic mode."""
    import numpy as np

    import fuzz_parity
    import gnuspeech_b200 as g
    from gnuspeech_b200 import workloads as W
    rng = np.random.default_rng(seed)
    ip = fuzz_parity.voice(rng)
    n, nf = 4, int(rng.integers(60, 220))
    frames = W.random_walk(n, nf, seed=300 + seed).reshape(n, nf, 16)
    for u in range(n):
        fuzz_parity.patch_edges(frames[u], rng)
    pushes = []
    while sum(pushes) < nf:
        pushes.append(int(min(nf - sum(pushes), rng.integers(1, 40))))
    b = g.TRMBatch(ip, [nf] * n, precision=precision)
    smp = np.zeros(b.layout.total_out_samples, b.sample_dtype)
    b.synthesize(frames.reshape(n * nf, 16), samples_out=smp, devices=[0])
    ns, oo = b.numberSamples, b.outOffsets
    st = g.TRMStream(n, ip, precision=precision, max_frames_per_push=max(pushes))
    got = [[] for _ in range(n)]
    at = 0
    for k, m in enumerate(pushes):
        out = st.push(frames[:, at:at + m], flush=(k == len(pushes) - 1))
        at += m
        for u in range(n):
            got[u].append(out[u])
    st.free()
    for u in range(n):
        y = np.concatenate(got[u])
        want = smp[oo[u]:oo[u] + ns[u]]
        assert y.shape == want.shape, (u, y.shape, want.shape)
        assert np.array_equal(y, want, equal_nan=True), "stream %d: %d samples differ" % (u, int((y != want).sum()))
