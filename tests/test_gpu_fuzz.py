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
