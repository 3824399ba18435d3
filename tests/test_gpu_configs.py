"""GPU parity on slices of every BASELINE.json config and on the edge cases the path has: ragged lengths, odd batch
sizes (half-empty warps), single-frame and empty utterances, mixed voices / output rates (up- and down-sampling),
stereo, the sine waveform, long static vowels (the worst case for an FP32 frequency path), and size-independent
properties at full size."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu


FP64_TOL = 1e-9
FP32_SNR_DB = 80.0


def _g():
    import gnuspeech_b200 as g
    return g


def _run(ips, frames, n_frames, precision, want_tube=False):
    g = _g()
    b = g.TRMBatch(ips, n_frames, precision=precision)
    lay = b.layout
    pcm = np.zeros(max(1, lay.total_pcm_samples), np.int16)
    smp = np.zeros(max(1, lay.total_out_samples), b.sample_dtype)
    tube = np.zeros(max(1, b.tubeElements), b.sample_dtype) if want_tube else None
    if want_tube:
        b.synthesize_debug(frames, pcm, smp, tube)
    else:
        b.synthesize(frames, pcm_out=pcm, samples_out=smp, devices=[0])
    return b, pcm, smp, tube


def _check(b, pcm, smp, ips, frames, n_frames, precision, idx=None):
    g = _g()
    ns, po, oo, mx = b.numberSamples, b.pcmOffsets, b.outOffsets, b.maximumSampleValues
    off = np.concatenate(([0], np.cumsum(n_frames)))
    is64 = precision != g.TRM_PRECISION_FP32
    worst = 0.0 if is64 else 1e9
    for u in (range(len(n_frames)) if idx is None else idx):
        ip = ips[u] if isinstance(ips, (list, tuple)) else ips
        ref = O.synthesize(ip, frames[off[u]:off[u + 1]], want_tube=False)
        assert ns[u] == ref.numberSamples, "utt %d" % u
        if ref.numberSamples == 0:
            continue
        ch = 2 if ip.channels == 2 else 1
        y = smp[oo[u]:oo[u] + ns[u]].astype(np.float64)
        p = pcm[po[u]:po[u] + ns[u] * ch].astype(np.int32)
        peak = ref.maximumSampleValue
        if peak == 0.0:
            assert not y.any() and not p.any()
            continue
        pcm_ref = O.pcm16(ip, ref.samples, peak).astype(np.int32)
        if is64:
            e = np.abs(y - ref.samples).max() / peak
            assert e <= FP64_TOL, "utt %d: FP64 error %.3e" % (u, e)
            assert abs(mx[u] - peak) <= FP64_TOL * peak
            worst = max(worst, e)
        else:
            snr = O.snr_db(ref.samples, y)
            assert snr >= FP32_SNR_DB, "utt %d: SNR %.1f dB" % (u, snr)
            worst = min(worst, snr)
        d = np.abs(p - pcm_ref)
        assert d.max() <= 1, "utt %d: %d PCM samples off by more than 1 LSB (max %d)" % (u, int((d > 1).sum()), int(d.max()))
    return worst


@pytest.mark.parametrize("precision", [0, 1, 2])
def test_config3_static_grid_slice(precision):
    """Config 3: TRAcT-style static sweep (0.5 s each): 96 grid points spread over the 65,536."""
    g = _g()
    from gnuspeech_b200 import workloads as W
    idx = [int(i) for i in np.linspace(0, 65535, 96)]
    nf = 126
    frames = W.grid(idx, nf)
    ip = g.TRMInputParameters(44100.0)
    b, pcm, smp, _ = _run(ip, frames, [nf] * len(idx), precision)
    assert (b.numberSamples == 22109).all()
    print("config3 worst", _check(b, pcm, smp, ip, frames, [nf] * len(idx), precision))


@pytest.mark.parametrize("precision", [0, 1, 2])
def test_config4_mixed_lengths_and_rates(precision):
    """Config 4 slice: ragged lengths, alternating 44.1 / 22.05 kHz, odd utterance count, plus the degenerate
    single-frame (flush only) and two-frame utterances."""
    g = _g()
    from gnuspeech_b200 import workloads as W
    n_frames = [751, 1, 313, 2, 1251, 126, 433, 877, 51]
    frames = W.random_walk_ragged(n_frames, seed=21)
    ips = [g.TRMInputParameters(44100.0 if u % 2 == 0 else 22050.0) for u in range(len(n_frames))]
    b, pcm, smp, _ = _run(ips, frames, n_frames, precision)
    print("config4 worst", _check(b, pcm, smp, ips, frames, n_frames, precision))


@pytest.mark.parametrize("precision", [0, 1, 2])
def test_mixed_voices_down_sampling_stereo_sine(precision):
    """Different tube lengths in one batch (different tube rates, control periods, converter directions and pads),
    stereo output with balance / volume, sine glottal source, modulation off, other pulse shapes."""
    g = _g()
    from gnuspeech_b200 import workloads as W
    ips = [
        g.TRMInputParameters(44100.0, length=15.0),
        g.TRMInputParameters(22050.0, length=15.0),                       # down-sampling, pad 14
        g.TRMInputParameters(22050.0, length=10.0, temperature=32.0),
        g.TRMInputParameters(44100.0, length=7.5),                        # down-sampling at 44.1 kHz
        g.TRMInputParameters(22050.0, channels=2, balance=-0.4, volume=57.0),
        g.TRMInputParameters(44100.0, waveform=1),
        g.TRMInputParameters(44100.0, usesModulation=0, breathiness=4.0, lossFactor=1.5),
        g.TRMInputParameters(44100.0, tp=30.0, tnMin=12.0, tnMax=40.0, mixOffset=48.0, throatVol=12.0),
        g.TRMInputParameters(22050.0, length=5.0),     # 70 kHz tube rate: ratio 0.315, pad 42 -- the converter window nearly
                                                       # fills the 128 staged rows (8 outputs per work item)
    ]
    n_frames = [101] * len(ips)
    frames = W.random_walk(len(ips), 101, seed=33)
    b, pcm, smp, _ = _run(ips, frames, n_frames, precision)
    print("mixed voices worst", _check(b, pcm, smp, ips, frames, n_frames, precision))


def test_long_static_vowel_fp32_phase_accuracy():
    """30 s static /aa/: an FP32 frequency path drifts here (19 dB at 10 s); the fast mode
    keeps pitch -> f0 -> phase in FP64 / 64-bit fixed point."""
    g = _g()
    from gnuspeech_b200 import workloads as W
    nf = 7501
    frames = W.static_vowel(nf, 1)
    ip = g.TRMInputParameters(44100.0)
    b, pcm, smp, _ = _run(ip, frames, [nf], g.TRM_PRECISION_FP32)
    print("30 s static vowel SNR", _check(b, pcm, smp, ip, frames, [nf], g.TRM_PRECISION_FP32))


def test_long_random_walk_both_modes():
    """One 60 s random-walk utterance (config 4's longest): 1.185e6 tube samples, 2.646e6 output samples."""
    g = _g()
    from gnuspeech_b200 import workloads as W
    nf = 15001
    frames = W.random_walk(1, nf, seed=44)
    ip = g.TRMInputParameters(44100.0)
    for prec in (g.TRM_PRECISION_FP64, g.TRM_PRECISION_FP32):
        b, pcm, smp, _ = _run(ip, frames, [nf], prec)
        assert b.numberSamples[0] == 2646061
        print("60 s walk, precision %d:" % prec, _check(b, pcm, smp, ip, frames, [nf], prec))


def test_noise_and_frame_indexing_are_exact():
    """North-star: the noise sequence and frame indexing must be bit-exact.  With everything but aspiration noise
    silenced (glottal volume 0, frication 0) the tube input is ah1 * lp_noise * 0.125: any slip in the jump-ahead
    MCG or in which frame a sample belongs to shows at full scale; FP64 must stay at rounding level."""
    g = _g()
    from gnuspeech_b200 import workloads as W
    nf = 201
    frames = W.static_vowel(nf, 0)
    frames[:, 1] = 0.0                      # no voicing
    frames[:, 2] = np.linspace(0.0, 40.0, nf)   # aspiration ramps up: frame indexing matters
    ip = g.TRMInputParameters(44100.0, usesModulation=0)
    ref = O.synthesize(ip, frames)
    b, pcm, smp, tube = _run(ip, frames, [nf], g.TRM_PRECISION_FP64_STRICT, want_tube=True)
    t = tube[:ref.tube.size]
    assert np.abs(t - ref.tube).max() <= 1e-13 * np.abs(ref.tube).max()
    # first samples: x[0] uses frame 0 exactly, the increment is applied AFTER each sample
    assert t[0] == ref.tube[0] and t[1] == ref.tube[1] and t[79] == ref.tube[79]
    # conformance mode: the same noise draws (integer generator) through fused arithmetic -- rounding level, not bit level
    b, pcm, smp, tube = _run(ip, frames, [nf], g.TRM_PRECISION_FP64, want_tube=True)
    assert np.abs(tube[:ref.tube.size] - ref.tube).max() <= 1e-12 * np.abs(ref.tube).max()


def test_config5_slice_properties_at_scale():
    """Config 5 / config 2 at scale (2048 x 2 s, FP32 fast): results do not depend on batch composition --
    every utterance of the big batch equals the same utterance synthesized alone -- and a sample of them matches
    the oracle."""
    g = _g()
    from gnuspeech_b200 import workloads as W
    n, nf = 2048, 501
    frames = W.random_walk(n, nf, seed=5)
    ip = g.TRMInputParameters(44100.0)
    b, pcm, smp, _ = _run(ip, frames, [nf] * n, g.TRM_PRECISION_FP32)
    ns, po = b.numberSamples, b.pcmOffsets
    assert (ns == 88259).all()
    pick = [0, 1, 511, 512, 1000, 2047]
    _check(b, pcm, smp, ip, frames, [nf] * n, g.TRM_PRECISION_FP32, idx=pick)
    for u in pick:
        b1, pcm1, smp1, _ = _run(ip, frames[u * nf:(u + 1) * nf], [nf], g.TRM_PRECISION_FP32)
        assert np.array_equal(pcm1[:ns[u]], pcm[po[u]:po[u] + ns[u]]), "utterance %d depends on its batch" % u
    # every utterance is normalised to its own peak: the loudest PCM code of each is +-32767 (volume 60 dB)
    peaks = np.array([np.abs(pcm[po[u]:po[u] + ns[u]].astype(np.int32)).max() for u in range(0, n, 37)])
    assert (peaks == 32767).all()


def test_multi_device_call_matches_single_device():
    """TRMBatchSynthesize over every visible GPU (one host thread and context per device, contiguous shards, no
    collectives) gives the same bytes as one device."""
    import torch
    g = _g()
    from gnuspeech_b200 import workloads as W
    n, nf = 37, 126
    frames = W.random_walk(n, nf, seed=8)
    ip = g.TRMInputParameters(44100.0)
    b, pcm, smp, _ = _run(ip, frames, [nf] * n, g.TRM_PRECISION_FP64)
    nd = torch.cuda.device_count()
    b2 = g.TRMBatch(ip, [nf] * n, precision=g.TRM_PRECISION_FP64)
    pcm2 = np.zeros_like(pcm)
    b2.synthesize(frames, pcm_out=pcm2, devices=list(range(nd)))
    ns, po = b.numberSamples, b.pcmOffsets
    for u in range(n):                                   # (the alignment padding between utterances is undefined)
        assert np.array_equal(pcm[po[u]:po[u] + ns[u]], pcm2[po[u]:po[u] + ns[u]]), u
    assert np.array_equal(b.maximumSampleValues, b2.maximumSampleValues)


def test_synthesizer_adapter_and_file_outputs(tmp_path):
    """The TRMSynthesizer call pattern Monet uses (TRMSynthesizer.m:38-136) and the three container writers."""
    g = _g()
    from gnuspeech_b200 import workloads as W
    syn = g.TRMSynthesizer()
    syn.setupSynthesisParameters(g.TRMInputParameters(22050.0))
    for f in W.static_vowel(26, 0):
        syn.addParameters(g.TRMParameters(*f[:7], radius=f[7:15], velum=f[15]))
    tube = syn.synthesize()
    ref = O.synthesize(g.TRMInputParameters(22050.0), W.static_vowel(26, 0))
    assert tube.numberSamples == ref.numberSamples
    assert syn.lastWAVData == O.wav_bytes(g.TRMInputParameters(22050.0), tube.resampledData, tube.maximumSampleValue)
    pcm = tube.pcm16()
    for fmt, magic in ((0, b".snd"), (1, b"FORM"), (2, b"RIFF")):
        syn2 = g.TRMSynthesizer()
        syn2.setupSynthesisParameters(g.TRMInputParameters(22050.0))
        syn2.fileType = fmt
        syn2.shouldSaveToSoundFile = True
        syn2.filename = str(tmp_path / ("out%d" % fmt))
        syn2.addParameters(W.static_vowel(26, 0))
        syn2.synthesize()
        raw = open(syn2.filename, "rb").read()
        assert raw[:4] == magic
        body = np.frombuffer(raw[-2 * pcm.size:], dtype="<i2" if fmt == 2 else ">i2")
        assert np.array_equal(body.astype(np.int16), pcm)


def test_async_calls_overlap_and_match_blocking():
    """TRMBatchSynthesizeAsync / TRMBatchWait: two calls in flight on one device (two context lanes) return the same
    bytes as the blocking call, in any completion order; a ticket reports errors like the blocking call does."""
    g = _g()
    from gnuspeech_b200 import workloads as W
    n, nf = 96, 126
    ip = g.TRMInputParameters(44100.0)
    frames_a = W.random_walk(n, nf, seed=41)
    frames_b = W.random_walk(n, nf, seed=42)
    ref = []
    for fr in (frames_a, frames_b):
        b, pcm, _, _ = _run(ip, fr, [nf] * n, g.TRM_PRECISION_FP64)
        ref.append((pcm.copy(), b.maximumSampleValues.copy()))
    ns, po = b.numberSamples, b.pcmOffsets
    batches = [g.TRMBatch(ip, [nf] * n, precision=g.TRM_PRECISION_FP64) for _ in range(2)]
    pcms = [np.zeros(batches[0].layout.total_pcm_samples, np.int16) for _ in range(2)]
    for rounds in range(3):
        tickets = [batches[k].synthesize_async((frames_a, frames_b)[k], pcm_out=pcms[k], devices=[0]) for k in range(2)]
        for k in (1, 0):
            tickets[k].wait()
        for k in range(2):
            assert np.array_equal(batches[k].maximumSampleValues, ref[k][1])
            for u in range(n):                               # (the alignment padding between utterances is undefined)
                assert np.array_equal(pcms[k][po[u]:po[u] + ns[u]], ref[k][0][po[u]:po[u] + ns[u]]), (rounds, k, u)
            pcms[k][:] = 0


def _spot_check(b, pcm, ips, frames, n_frames, precision, picks):
    """oracle comparison of a few utterances of a big batch: PCM within 1 LSB, numberSamples, maxima"""
    g = _g()
    ns, po, mx = b.numberSamples, b.pcmOffsets, b.maximumSampleValues
    off = np.concatenate(([0], np.cumsum(n_frames)))
    for u in picks:
        ip = ips[u] if isinstance(ips, (list, tuple)) else ips
        ref = O.synthesize(ip, frames[off[u]:off[u + 1]], want_tube=False)
        assert ns[u] == ref.numberSamples, u
        tol = 1e-9 if precision != g.TRM_PRECISION_FP32 else 2e-5
        assert abs(mx[u] - ref.maximumSampleValue) <= tol * ref.maximumSampleValue, u
        pcm_ref = O.pcm16(ip, ref.samples, ref.maximumSampleValue).astype(np.int32)
        d = np.abs(pcm[po[u]:po[u] + ns[u]].astype(np.int32) - pcm_ref)
        assert d.max() <= 1, "utterance %d: PCM off by %d LSB" % (u, int(d.max()))


def test_config2_full_size_fp32():
    """configs[1] at BASELINE size: 4096 random-walk utterances x 10 s, FP32 fast mode, through TRMBatchSynthesize with
    host buffers.  Properties that do not need the oracle at this size: every utterance has the sample count of the
    closed form and is normalised to full scale (its loudest PCM code is +-32767 at 60 dB volume); a spread of
    utterances is compared with the oracle (+-1 LSB)."""
    g = _g()
    from gnuspeech_b200 import workloads as W
    n, nf = 4096, 2501
    frames = g.PinnedArray((n * nf, 16), np.float64)
    W.random_walk(n, nf, seed=1, out=frames.array)
    ip = g.TRMInputParameters(44100.0)
    b = g.TRMBatch(ip, [nf] * n, precision=g.TRM_PRECISION_FP32)
    pcm = g.PinnedArray(int(b.layout.total_pcm_samples), np.int16)
    b.synthesize(frames, pcm_out=pcm, devices=[0])
    ns, po = b.numberSamples, b.pcmOffsets
    assert (ns == 441059).all()
    peaks = np.array([np.abs(pcm.array[po[u]:po[u] + ns[u]]).max() for u in range(0, n, 16)])
    assert (peaks == 32767).all()
    assert np.isfinite(b.maximumSampleValues).all() and (b.maximumSampleValues > 0).all()
    _spot_check(b, pcm.array, ip, frames.array, [nf] * n, g.TRM_PRECISION_FP32, [0, 1337, 4095])
    pcm.free()
    frames.free()


def test_config2_full_size_fp64_spot_check():
    """configs[1] at BASELINE size in the headline precision: 4096 random-walk utterances x 10 s, FP64 conformance, through
    TRMBatchSynthesize with host buffers (the path that uses the time split and the output groups).  First, middle and
    last utterance plus two others against the oracle: 1e-9 on the output samples, +-1 LSB on the PCM."""
    g = _g()
    from gnuspeech_b200 import workloads as W
    n, nf = 4096, 2501
    frames = g.PinnedArray((n * nf, 16), np.float64)
    W.random_walk(n, nf, seed=1, out=frames.array)
    ip = g.TRMInputParameters(44100.0)
    b = g.TRMBatch(ip, [nf] * n, precision=g.TRM_PRECISION_FP64)
    pcm = g.PinnedArray(int(b.layout.total_pcm_samples), np.int16)
    b.synthesize(frames, pcm_out=pcm, devices=[0])
    ns, po = b.numberSamples, b.pcmOffsets
    assert (ns == 441059).all()
    assert np.isfinite(b.maximumSampleValues).all() and (b.maximumSampleValues > 0).all()
    _spot_check(b, pcm.array, ip, frames.array, [nf] * n, g.TRM_PRECISION_FP64, [0, 1337, 2048, 3000, 4095])
    pcm.free()
    frames.free()


def test_config3_full_grid_fp32():
    """configs[2] at BASELINE size: the full 65,536-point static grid x 0.5 s (TRAcT-style sweep), FP32 fast mode.
    Static vowels: the two pitches of the grid give identical spectra up to the source, the closed-form sample count
    holds for every utterance, every utterance is normalised to full scale; grid corners are compared with the oracle."""
    g = _g()
    from gnuspeech_b200 import workloads as W
    n, nf = 65536, 126
    frames = W.grid(range(n), nf)
    ip = g.TRMInputParameters(44100.0)
    b = g.TRMBatch(ip, [nf] * n, precision=g.TRM_PRECISION_FP32)
    pcm = np.zeros(int(b.layout.total_pcm_samples), np.int16)
    b.synthesize(frames, pcm_out=pcm, devices=[0])
    ns, po = b.numberSamples, b.pcmOffsets
    assert (ns == 22109).all()
    peaks = np.array([np.abs(pcm[po[u]:po[u] + ns[u]]).max() for u in range(0, n, 97)])
    assert (peaks == 32767).all()
    # same tract, same pitch, different batch position -> same bytes (grid index bit 14 = velum, bit 15 = pitch)
    b1 = g.TRMBatch(ip, [nf], precision=g.TRM_PRECISION_FP32)
    for u in (0, 12345, 65535):
        one = np.zeros(int(b1.layout.total_pcm_samples), np.int16)
        b1.synthesize(frames[u * nf:(u + 1) * nf], pcm_out=one, devices=[0])
        d = np.abs(one[:ns[u]].astype(np.int32) - pcm[po[u]:po[u] + ns[u]].astype(np.int32))
        assert d.max() <= 1, u                            # +-1 LSB: the lone utterance runs the other lane mapping
    _spot_check(b, pcm, ip, frames, [nf] * n, g.TRM_PRECISION_FP32, [0, 21845, 43690, 65535])


def test_config4_full_size_mixed_lengths():
    """configs[3] at BASELINE size: 256 utterances of 5-60 s, alternating 44.1 / 22.05 kHz (load balancing: longest
    first inside the kernels, ragged everywhere), FP32 fast mode; the shortest, the longest and two others are compared
    with the oracle, all are checked for sample count and normalisation."""
    g = _g()
    from gnuspeech_b200 import workloads as W
    rng = np.random.default_rng(4)
    n = 256
    n_frames = [int(x) for x in rng.integers(5 * 250, 60 * 250 + 1, n)]
    n_frames[7], n_frames[200] = 5 * 250 + 1, 60 * 250 + 1
    frames = W.random_walk_ragged(n_frames, seed=9)
    ips = [g.TRMInputParameters(44100.0 if u % 2 == 0 else 22050.0) for u in range(n)]
    b = g.TRMBatch(ips, n_frames, precision=g.TRM_PRECISION_FP32)
    pcm = np.zeros(int(b.layout.total_pcm_samples), np.int16)
    b.synthesize(frames, pcm_out=pcm, devices=[0])
    ns, po = b.numberSamples, b.pcmOffsets
    for u in range(n):
        want = g.derive(ips[u], n_frames[u]).numberSamples
        assert ns[u] == want, u
        assert np.abs(pcm[po[u]:po[u] + ns[u]]).max() == 32767, u
    _spot_check(b, pcm, ips, frames, n_frames, g.TRM_PRECISION_FP32, [7, 200, 64, 129])


@pytest.mark.parametrize("precision", [0, 1, 2])
def test_time_split_waveguide_is_bit_identical(precision, monkeypatch):
    """Long chunks whose utterances have equal frame counts run the waveguide as two launches in time (the later frames
    are uploaded behind the first launch; recurrence state carried like a streaming push, restart inside a control
    interval).  Mixed voices give every utterance its own split point and restart phase.  The result must equal the
    single-launch result bit for bit in both precision modes, and the oracle within tolerance."""
    g = _g()
    from gnuspeech_b200 import workloads as W
    # one tube length (one control period: every utterance reaches the split after the same number of blocks), otherwise mixed
    voices = [dict(), dict(waveform=1), dict(usesModulation=0, breathiness=3.0), dict(channels=2, balance=0.3), dict(lossFactor=1.2)]
    n, nf = 37, 600
    ips = [g.TRMInputParameters(44100.0 if u % 4 else 22050.0, **voices[u % len(voices)]) for u in range(n)]
    frames = W.random_walk(n, nf, seed=41)
    monkeypatch.delenv("TRM_NO_TIME_SPLIT", raising=False)
    b1, pcm1, smp1, _ = _run(ips, frames, [nf] * n, precision)
    assert b1.kernelLaunches == 4                      # two waveguide launches + resampler + PCM
    monkeypatch.setenv("TRM_NO_TIME_SPLIT", "1")
    b0, pcm0, smp0, _ = _run(ips, frames, [nf] * n, precision)
    assert b0.kernelLaunches == 3
    for u in range(n):
        o, k = b0.outOffsets[u], b0.numberSamples[u]
        assert np.array_equal(smp1[o:o + k], smp0[o:o + k]), u
        ch = ips[u].channels
        assert np.array_equal(pcm1[b0.pcmOffsets[u]:b0.pcmOffsets[u] + k * ch], pcm0[b0.pcmOffsets[u]:b0.pcmOffsets[u] + k * ch]), u
    assert np.array_equal(b1.maximumSampleValues, b0.maximumSampleValues)
    print("time split worst", _check(b1, pcm1, smp1, ips, frames, [nf] * n, precision, idx=[0, 1, 2, 36]))
    # different tube lengths in one chunk: no split
    monkeypatch.delenv("TRM_NO_TIME_SPLIT", raising=False)
    ips[5] = g.TRMInputParameters(44100.0, length=15.0)
    b2, _, _, _ = _run(ips, frames, [nf] * n, precision)
    assert b2.kernelLaunches == 3


def test_output_groups_equal_single_pass(monkeypatch):
    """Chunks with more than 128 M output samples are resampled and scaled in up to four output groups whose PCM leaves
    while the next group is resampled.  Ragged lengths, two output rates and three tube lengths, so the groups cut
    through converter signatures: PCM and maxima must equal the single-pass result exactly."""
    g = _g()
    from gnuspeech_b200 import workloads as W
    n = 4000
    rng = np.random.default_rng(77)
    n_frames = [int(v) for v in rng.integers(150, 600, n)]
    voices = [dict(), dict(length=15.0), dict(length=10.0, temperature=32.0)]
    ips = [g.TRMInputParameters(44100.0 if u % 3 else 22050.0, **voices[(u // 7) % 3]) for u in range(n)]
    frames = np.concatenate([W.random_walk(1, nf, seed=1000 + u) for u, nf in enumerate(n_frames)])
    res = []
    for off in (False, True):
        if off:
            monkeypatch.setenv("TRM_NO_OUT_GROUPS", "1")
        else:
            monkeypatch.delenv("TRM_NO_OUT_GROUPS", raising=False)
        b = g.TRMBatch(ips, n_frames, precision=g.TRM_PRECISION_FP64)      # (FP64: the whole batch is one chunk)
        pcm = np.zeros(b.layout.total_pcm_samples, np.int16)
        b.synthesize(frames, pcm_out=pcm, devices=[0])
        res.append((b, pcm))
    (b1, p1), (b0, p0) = res
    assert b0.kernelLaunches == 3 and b1.kernelLaunches in (5, 7, 9)        # waveguide + (resampler + PCM) per group
    assert np.array_equal(b1.maximumSampleValues, b0.maximumSampleValues)
    po, ns = b0.pcmOffsets, b0.numberSamples
    for u in range(n):
        assert np.array_equal(p1[po[u]:po[u] + ns[u]], p0[po[u]:po[u] + ns[u]]), u
