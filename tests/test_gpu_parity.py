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


@pytest.fixture(params=["sections", "utterances"], autouse=True)
def tube_mapping(request, monkeypatch):
    """Every test runs with both waveguide mappings: lane-per-section (tube_kernel.cuh, what small batches get by
    default) and the batch-throughput mapping (tube_wide.cuh, what large batches get)."""
    monkeypatch.setenv("TRM_TUBE_MAPPING", request.param)
    return request.param

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


@pytest.mark.parametrize("posture", [0, 1])
@pytest.mark.parametrize("rate", [44100.0, 22050.0])
def test_config1_static_vowel_fp64(posture, rate):
    """Config 1: 1 s static vowel, male voice, 250 Hz control frames."""
    g = _g()
    from gnuspeech_b200 import workloads as W
    ip = g.TRMInputParameters(rate)
    frames = W.static_vowel(251, posture)
    ref = O.synthesize(ip, frames)
    m = _model(ip, frames, g.TRM_PRECISION_FP64)
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


def test_fixture_gnuspeech_input_fp64():
    """The one real TRM input file the reference ships (Applications/Monet/samples/gnuspeech.input)."""
    g = _g()
    path = os.path.join(GOLDEN, "gnuspeech.input")
    oip, oframes = O.parse_input_file(path)
    dl = g.TRMDataList(path)
    assert dl.count == oframes.shape[0] == 344
    assert np.array_equal(dl.values, oframes)
    ref = O.synthesize(oip, oframes)
    m = g.TRMTubeModel(dl)
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


def test_config2_random_walk_batch_fp64():
    """Config 2 slice: 12 random-walk utterances x 2 s, all parameters varying (incl. frication, velum)."""
    g = _g()
    from gnuspeech_b200 import workloads as W
    ip = g.TRMInputParameters(44100.0)
    n, nf = 12, 501
    frames = W.random_walk(n, nf, seed=2)
    worst = 0.0
    for u, (ref, y, t, p, mx) in enumerate(_batch_vs_oracle(ip, frames, [nf] * n, g.TRM_PRECISION_FP64)):
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
