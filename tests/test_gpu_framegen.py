"""GPU parity of the control-frame generator (framegen_kernel.cuh; EventList.m:883-1061 + MMDriftGenerator.m) against the
oracle's restatement: every frame value and the drift generator's final seed are identical bits; the fused
event-lists -> PCM path gives the bytes of the frames -> PCM path."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu


def _g():
    import gnuspeech_b200 as g
    return g


def _lists(n, seconds, smooth):
    evs = [O.synthetic_event_list(1000 + 7 * u, seconds[u % len(seconds)], smooth=smooth) for u in range(n)]
    counts = [len(e) for e in evs]
    return evs, counts, np.concatenate(evs)


@pytest.mark.parametrize("mode", ["smooth+drift", "plain", "macro only", "per-utterance"])
def test_generated_frames_equal_the_oracle(mode):
    g = _g()
    n = 37
    smooth = mode in ("smooth+drift", "per-utterance")
    evs, counts, allev = _lists(n, [0.6, 1.3, 2.1], smooth)
    if mode == "smooth+drift":
        fgs = g.TRMFrameGeneration()
    elif mode == "plain":
        fgs = g.TRMFrameGeneration(useDrift=0, useMacroIntonation=0, useSmoothIntonation=0, useMicroIntonation=1, pitch=-3.5)
    elif mode == "macro only":
        fgs = g.TRMFrameGeneration(useDrift=1, useMacroIntonation=1, useSmoothIntonation=0, useMicroIntonation=0, driftDeviation=0.5,
                                   driftCutoff=200.0)          # cutoff above sampleRate/2 is clamped (MMDriftGenerator.m:46-48)
    else:
        fgs = [g.TRMFrameGeneration(useDrift=u % 2, useSmoothIntonation=(u // 2) % 2, useMacroIntonation=(u // 4) % 2, pitch=-12.0 + u,
                                    driftSeed=np.float32(0.1 + 0.02 * u)) for u in range(n)]
    n_frames = [g.event_list_frame_count(e) for e in evs]
    assert all(nf == O.frame_count(e) for nf, e in zip(n_frames, evs))
    ip = g.TRMInputParameters(44100.0)
    b = g.TRMBatch(ip, n_frames, precision=g.TRM_PRECISION_FP64)
    frames, seeds = b.generate_frames(allev, counts, fgs)
    off = np.concatenate(([0], np.cumsum(n_frames)))
    for u in range(n):
        ref, seed = O.generate_frames(evs[u], fgs[u] if isinstance(fgs, list) else fgs)
        got = frames[off[u]:off[u + 1]]
        assert got.shape == ref.shape
        assert np.array_equal(got, ref), "utterance %d: %d of %d values differ" % (u, int((got != ref).sum()), ref.size)
        assert seeds[u] == np.float32(seed), u


def test_events_to_pcm_equals_frames_to_pcm():
    """TRMBatchSynthesizeEvents (generator + tube model on the device, frames never on the host) == TRMBatchSynthesize on
    the frames TRMBatchGenerateFrames returns, for both precisions; and a frame-count mismatch is refused."""
    g = _g()
    n = 24
    evs, counts, allev = _lists(n, [0.4, 0.9], True)
    n_frames = [g.event_list_frame_count(e) for e in evs]
    ip = g.TRMInputParameters(44100.0)
    fg = g.TRMFrameGeneration()
    for prec in (g.TRM_PRECISION_FP64, g.TRM_PRECISION_FP32):
        b = g.TRMBatch(ip, n_frames, precision=prec)
        frames, _ = b.generate_frames(allev, counts, fg)
        pcm_a = np.zeros(b.layout.total_pcm_samples, np.int16)
        pcm_b = np.zeros_like(pcm_a)
        b.synthesize(frames, pcm_out=pcm_a, devices=[0])
        max_a = b.maximumSampleValues.copy()
        b.synthesize_events(allev, counts, fg, pcm_out=pcm_b)
        assert np.array_equal(b.maximumSampleValues, max_a)
        ns, po = b.numberSamples, b.pcmOffsets
        for u in range(n):
            assert np.array_equal(pcm_a[po[u]:po[u] + ns[u]], pcm_b[po[u]:po[u] + ns[u]]), u
        assert np.abs(pcm_a).max() > 1000
    bad = g.TRMBatch(ip, [nf + 1 for nf in n_frames], precision=g.TRM_PRECISION_FP32)
    with pytest.raises(g.TRMError):
        bad.synthesize_events(allev, counts, fg, pcm_out=np.zeros(bad.layout.total_pcm_samples, np.int16))
