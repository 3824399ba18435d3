"""GPU parity tests: the CUDA path (through the C-ABI, libtrm -> libtrm_cuda) against the CPU oracle on the
same inputs.  Tolerances are BASELINE.json's:
  FP64 conformance : max|y_gpu - y_ref| <= 1e-9 * max|y_ref| per utterance, numberSamples equal, max within 1e-9
  FP32 fast        : SNR >= 80 dB per utterance on output-rate samples and |pcm - pcm_ref| <= 1 LSB
"""
import os

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu


# FP64 modes of the GPU path: conformance (cheaper arithmetic forms, contract 1e-9) and strict (the reference's operations
# in the reference's order; the bit-faithful twin).  Both must meet the same contract against the oracle.
FP64_MODES = [0, 2]

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FP64_TOL = 1e-9
FP32_SNR_DB = 80.0


def _g():
    import gnuspeech_b200 as g
    return g


def _model(ip, frames, precision):
    g = _g()
    dl = g.TRMDataList()
    dl.setInputParameters(ip)
    dl.addParameters(np.ascontiguousarray(frames, dtype=np.float64))
    m = g.TRMTubeModel(dl, precision=precision)
    m.synthesize()
    return m


def _check_fp64(m, ref, what):
    assert m.numberSamples == ref.numberSamples, what
    peak_t = np.abs(ref.tube).max() if ref.tube.size else 0.0
    if ref.tube.size:
        et = np.abs(m.tubeSignal - ref.tube).max()
        assert et <= FP64_TOL * peak_t, "%s: tube-rate error %.3e of peak" % (what, et / peak_t)
    peak = ref.maximumSampleValue
    e = np.abs(m.resampledData - ref.samples).max() if ref.samples.size else 0.0
    assert e <= FP64_TOL * peak, "%s: output error %.3e of peak" % (what, e / max(peak, 1e-300))
    assert abs(m.maximumSampleValue - peak) <= FP64_TOL * peak, what
    return (e / peak) if peak else 0.0


def _check_fp32(m, ref, ip, what):
    assert m.numberSamples == ref.numberSamples, what
    snr = O.snr_db(ref.samples, m.resampledData)
    assert snr >= FP32_SNR_DB, "%s: SNR %.1f dB" % (what, snr)
    pcm_ref = O.pcm16(ip, ref.samples, ref.maximumSampleValue).astype(np.int32)
    pcm = m.pcm16().astype(np.int32)
    d = np.abs(pcm - pcm_ref)
    assert d.max() <= 1, "%s: %d samples off by more than 1 LSB (max %d)" % (what, int((d > 1).sum()), int(d.max()))
    return snr


@pytest.mark.parametrize("mode", FP64_MODES)
@pytest.mark.parametrize("posture", [0, 1])
@pytest.mark.parametrize("rate", [44100.0, 22050.0])
def test_config1_static_vowel_fp64(posture, rate, mode):
    """Config 1: 1 s static vowel, male voice, 250 Hz control frames."""
    g = _g()
    from gnuspeech_b200 import workloads as W
    ip = g.TRMInputParameters(rate)
    frames = W.static_vowel(251, posture)
    ref = O.synthesize(ip, frames)
    m = _model(ip, frames, mode)
    assert m.derived.controlPeriod == 79 and m.derived.sampleRate == 19750
    assert m.numberSamples == (44159 if rate == 44100.0 else 22080)
    _check_fp64(m, ref, "static vowel %d @%g" % (posture, rate))
    # PCM: bit-exact when both sides scale by their own max and the max agree to the last bit is not
    # guaranteed (libm differences); +-1 LSB is the contract, in practice FP64 is exact
    pcm_ref = O.pcm16(ip, ref.samples, ref.maximumSampleValue).astype(np.int32)
    assert np.abs(m.pcm16().astype(np.int32) - pcm_ref).max() <= 1


@pytest.mark.parametrize("posture", [0, 1])
def test_config1_static_vowel_fp32(posture):
    g = _g()
    from gnuspeech_b200 import workloads as W
    ip = g.TRMInputParameters(44100.0)
    frames = W.static_vowel(251, posture)
    ref = O.synthesize(ip, frames)
    m = _model(ip, frames, g.TRM_PRECISION_FP32)
    _check_fp32(m, ref, ip, "static vowel %d fp32" % posture)


@pytest.mark.parametrize("mode", FP64_MODES)
def test_fixture_gnuspeech_input_fp64(mode):
    """The one real TRM input file the reference ships (Applications/Monet/samples/gnuspeech.input)."""
    g = _g()
    path = os.path.join(GOLDEN, "gnuspeech.input")
    oip, oframes = O.parse_input_file(path)
    dl = g.TRMDataList(path)
    assert dl.count == oframes.shape[0] == 344
    assert np.array_equal(dl.values, oframes)
    ref = O.synthesize(oip, oframes)
    m = g.TRMTubeModel(dl, precision=mode)
    m.synthesize()
    _check_fp64(m, ref, "gnuspeech.input")
    assert m.generateWAVData() == O.wav_bytes(oip, m.resampledData, m.maximumSampleValue)


def test_fixture_gnuspeech_input_fp32():
    g = _g()
    path = os.path.join(GOLDEN, "gnuspeech.input")
    oip, oframes = O.parse_input_file(path)
    dl = g.TRMDataList(path)
    ref = O.synthesize(oip, oframes)
    m = g.TRMTubeModel(dl, precision=g.TRM_PRECISION_FP32)
    m.synthesize()
    _check_fp32(m, ref, oip, "gnuspeech.input fp32")


def _batch_vs_oracle(ip, frames, n_frames, precision, want_tube=True):
    g = _g()
    b = g.TRMBatch(ip, n_frames, precision=precision)
    lay = b.layout
    dt = b.sample_dtype
    pcm = np.zeros(max(1, lay.total_pcm_samples), np.int16)
    smp = np.zeros(max(1, lay.total_out_samples), dt)
    tube = np.zeros(max(1, b.tubeElements), dt)
    b.synthesize_debug(frames, pcm, smp, tube if want_tube else None)
    ns, po, oo, to, mx = b.numberSamples, b.pcmOffsets, b.outOffsets, b.tubeOffsets, b.maximumSampleValues
    at = 0
    results = []
    for u, nf in enumerate(n_frames):
        fr = frames[at:at + nf]
        at += nf
        ref = O.synthesize(ip, fr)
        assert ns[u] == ref.numberSamples
        y = smp[oo[u]:oo[u] + ns[u]].astype(np.float64)
        t = tube[to[u]:to[u] + ref.tube.size].astype(np.float64)
        p = pcm[po[u]:po[u] + ns[u]].astype(np.int32)
        results.append((ref, y, t, p, mx[u]))
    return results


@pytest.mark.parametrize("mode", FP64_MODES)
def test_config2_random_walk_batch_fp64(mode):
    """Config 2 slice: 12 random-walk utterances x 2 s, all parameters varying (incl. frication, velum)."""
    g = _g()
    from gnuspeech_b200 import workloads as W
    ip = g.TRMInputParameters(44100.0)
    n, nf = 12, 501
    frames = W.random_walk(n, nf, seed=2)
    worst = 0.0
    for u, (ref, y, t, p, mx) in enumerate(_batch_vs_oracle(ip, frames, [nf] * n, mode)):
        peak_t, peak = np.abs(ref.tube).max(), ref.maximumSampleValue
        et, e = np.abs(t - ref.tube).max() / peak_t, np.abs(y - ref.samples).max() / peak
        assert et <= FP64_TOL, "utt %d tube error %.3e" % (u, et)
        assert e <= FP64_TOL, "utt %d output error %.3e" % (u, e)
        assert abs(mx - peak) <= FP64_TOL * peak
        worst = max(worst, e)
        pcm_ref = O.pcm16(ip, ref.samples, ref.maximumSampleValue).astype(np.int32)
        assert np.abs(p - pcm_ref).max() <= 1
    print("worst FP64 relative error %.3e" % worst)


def test_config2_random_walk_batch_fp32():
    g = _g()
    from gnuspeech_b200 import workloads as W
    ip = g.TRMInputParameters(44100.0)
    n, nf = 12, 501
    frames = W.random_walk(n, nf, seed=2)
    worst = 1e9
    for u, (ref, y, t, p, mx) in enumerate(_batch_vs_oracle(ip, frames, [nf] * n, g.TRM_PRECISION_FP32)):
        snr = O.snr_db(ref.samples, y)
        assert snr >= FP32_SNR_DB, "utt %d SNR %.1f dB" % (u, snr)
        pcm_ref = O.pcm16(ip, ref.samples, ref.maximumSampleValue).astype(np.int32)
        d = np.abs(p - pcm_ref)
        assert d.max() <= 1, "utt %d: %d samples > 1 LSB (max %d)" % (u, int((d > 1).sum()), int(d.max()))
        worst = min(worst, snr)
    print("worst FP32 SNR %.1f dB" % worst)


# ---- edge cases the reference defines by what its code does (missing 5 and 7) ---------------------------

@pytest.mark.parametrize("precision", [0, 1, 2])
def test_adjacent_closed_sections_propagate_nan(precision):
    """Two adjacent sections with radius 0 make the junction coefficient 0/0 (TRMTubeModel.m:716-718).  The reference does
    not guard it: NaN enters the tube at that frame and never leaves (reproduce, do not fix).  The running
    maximum ignores NaN (`if (abs > max)`, TRMSampleRateConverter.m:206-208), so maximumSampleValue is the peak of the
    finite prefix; PCM of a NaN sample is 0.  Same NaN positions, same maximum, same PCM as the oracle."""
    g = _g()
    from gnuspeech_b200 import workloads as W
    nf = 101
    frames = W.static_vowel(nf, 1)
    frames[40:, 9] = 0.0          # radius[2] and radius[3] reach exactly 0 from frame 40 on (ramping down during interval 39)
    frames[40:, 10] = 0.0
    ip = g.TRMInputParameters(44100.0)
    ref = O.synthesize(ip, frames)
    assert np.isnan(ref.samples).any() and np.isfinite(ref.samples[:1000]).all() and ref.maximumSampleValue > 0
    (_, y, t, p, mx), = _batch_vs_oracle(ip, frames, [nf], precision)
    nan_ref = np.isnan(ref.samples)
    assert np.array_equal(np.isnan(y), nan_ref), "NaN does not start where the reference's does"
    assert np.array_equal(np.isnan(t), np.isnan(ref.tube))
    fin = ~nan_ref
    if precision == 1:
        assert O.snr_db(ref.samples[fin], y[fin]) >= FP32_SNR_DB
        assert abs(mx - ref.maximumSampleValue) <= 2e-5 * ref.maximumSampleValue
    else:
        assert np.abs(y[fin] - ref.samples[fin]).max() <= FP64_TOL * ref.maximumSampleValue
        assert abs(mx - ref.maximumSampleValue) <= FP64_TOL * ref.maximumSampleValue
    pcm_ref = O.pcm16(ip, ref.samples, ref.maximumSampleValue).astype(np.int32)
    assert (pcm_ref[nan_ref] == 0).all() and (p[nan_ref] == 0).all()
    assert np.abs(p - pcm_ref).max() <= 1


@pytest.mark.parametrize("precision", [0, 1, 2])
def test_frication_position_outside_the_taps(precision):
    """setFricationTaps (TRMTubeModel.m:748-765) truncates the position towards zero and compares an unsigned loop index
    with it: -0.5 -> tap FC1 gets 1.5 x amplitude and FC2 gets -0.5 x, below -1 nothing is injected, 7.4 feeds only the
    last tap (the second one would be past the array), 8 and above nothing."""
    g = _g()
    from gnuspeech_b200 import workloads as W
    nf = 151
    frames = W.static_vowel(nf, 0)
    frames[:, 1] = 0.0                                   # no voicing: frication noise only
    frames[:, 3] = 40.0                                  # frication volume
    pos = np.concatenate([np.linspace(-1.6, 0.6, 50), np.linspace(6.6, 8.3, 50), np.full(nf - 100, -0.5)])
    frames[:, 4] = pos.astype(np.float32)
    ip = g.TRMInputParameters(44100.0)
    (ref, y, t, p, mx), = _batch_vs_oracle(ip, frames, [nf], precision)
    assert ref.maximumSampleValue > 0
    cp = 79
    assert not ref.tube[: 10 * cp].any() and ref.tube[60 * cp: 70 * cp].any()      # below -1: silence; around 7: noise
    if precision == 1:
        assert O.snr_db(ref.samples, y) >= FP32_SNR_DB
    else:
        assert np.abs(t - ref.tube).max() <= FP64_TOL * np.abs(ref.tube).max()
        assert np.abs(y - ref.samples).max() <= FP64_TOL * ref.maximumSampleValue
    assert np.abs(p - O.pcm16(ip, ref.samples, ref.maximumSampleValue).astype(np.int32)).max() <= 1


def test_conformance_and_strict_modes_agree_at_scale():
    """FP64 conformance against its bit-faithful twin where the CPU oracle is too slow to be the judge: 1184 random-walk
    utterances x 4 s with mixed voices -- every sample within 1e-10 of the utterance's peak (contract 1e-9)."""
    g = _g()
    from gnuspeech_b200 import workloads as W
    n, nf = 1184, 1001
    frames = W.random_walk(n, nf, seed=31)
    voices = [dict(), dict(length=15.0), dict(waveform=1), dict(usesModulation=0, lossFactor=1.5), dict(length=12.0, temperature=30.0)]
    ips = [g.TRMInputParameters(44100.0 if u % 3 else 22050.0, **voices[u % len(voices)]) for u in range(n)]
    out = {}
    for mode in FP64_MODES:
        b = g.TRMBatch(ips, [nf] * n, precision=mode)
        smp = np.zeros(b.layout.total_out_samples, np.float64)
        b.synthesize(frames, samples_out=smp, devices=[0])
        out[mode] = (b, smp)
    (b0, s0), (b2, s2) = out[0], out[2]
    assert np.array_equal(b0.numberSamples, b2.numberSamples)
    worst = 0.0
    for u in range(n):
        o, k = b0.outOffsets[u], b0.numberSamples[u]
        worst = max(worst, np.abs(s0[o:o + k] - s2[o:o + k]).max() / b2.maximumSampleValues[u])
    print("conformance vs strict, worst of %d utterances: %.2e" % (n, worst))
    assert worst <= 1e-10
    assert np.abs(b0.maximumSampleValues - b2.maximumSampleValues).max() <= 1e-10 * b2.maximumSampleValues.max()


def test_long_static_vowel_fp64_phase():
    """30 s of constant pitch: the reference's double accumulator rounds every phase increment the same way each period and
    drifts from the exact phase (5.6e-11 of peak per second); the conformance mode's fixed-point phase emulates that
    rounding when the increment is constant (tube_wide.cuh) and must stay inside the contract."""
    g = _g()
    from gnuspeech_b200 import workloads as W
    nf = 7501
    frames = W.static_vowel(nf, 1)
    ip = g.TRMInputParameters(44100.0)
    ref = O.synthesize(ip, frames, want_tube=False)
    for mode, tol in ((0, 1e-10), (2, 1e-13)):
        b = g.TRMBatch(ip, [nf], precision=mode)
        smp = np.zeros(b.layout.total_out_samples, np.float64)
        b.synthesize(frames, samples_out=smp, devices=[0])
        e = np.abs(smp[:ref.numberSamples] - ref.samples).max() / ref.maximumSampleValue
        print("30 s static vowel, mode %d: %.2e" % (mode, e))
        assert e <= tol
