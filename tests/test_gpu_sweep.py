"""BASELINE configs[4] machinery: control tracks generated on the device, audio reduced to
per-utterance checksums on the device, shards of the index range independent of each other.  Plus the float32 frame
format and the copy probe of the end-to-end measurement."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu


def _g():
    import gnuspeech_b200 as g
    return g


@pytest.mark.parametrize("precision", [0, 1])
def test_sweep_equals_host_generated_batch(precision):
    """The device generator is the bit-for-bit twin of TRMWorkloadWalk2 and the checksum is what it says: a sweep over n
    utterances returns, for every utterance, the checksum of exactly the PCM that TRMBatchSynthesize produces from the
    host-generated tracks; probed utterances come back byte for byte."""
    g = _g()
    from gnuspeech_b200 import workloads as W
    n, nf, seed, first = 300, 126, 11, 1000
    ip = g.TRMInputParameters(44100.0)
    frames = W.walk2(n, nf, seed=seed, first_index=first)
    b = g.TRMBatch(ip, [nf] * n, precision=precision)
    pcm = np.zeros(b.layout.total_pcm_samples, np.int16)
    b.synthesize(frames, pcm_out=pcm, devices=[0])
    probes = [0, 7, 150, 299]
    r = g.sweep_synthesize(ip, nf, n, seed=seed, first_index=first, precision=precision, probes=probes)
    ns, po = b.numberSamples, b.pcmOffsets
    assert r["numberSamples"] == ns[0]
    want = np.array([g.pcm_checksum(pcm[po[u]:po[u] + ns[u]]) for u in range(n)], dtype=np.uint64)
    assert np.array_equal(r["checksums"], want)
    assert np.array_equal(r["maxima"], b.maximumSampleValues)
    for k, u in enumerate(probes):
        assert np.array_equal(r["probe_pcm"][k][:ns[u]], pcm[po[u]:po[u] + ns[u]]), u
    # and the tracks are a sensible workload: compare two probes with the oracle
    for k, u in list(enumerate(probes))[:2]:
        ref = O.synthesize(ip, frames[u * nf:(u + 1) * nf], want_tube=False)
        pcm_ref = O.pcm16(ip, ref.samples, ref.maximumSampleValue).astype(np.int32)
        assert np.abs(r["probe_pcm"][k][:ns[u]].astype(np.int32) - pcm_ref).max() <= 1


def test_sweep_shards_are_independent():
    """Any split of the index range gives the same per-utterance results (what sharding over GPUs relies on), also across
    the chunk boundary of one full wave (148 x 28 utterances) and with a short last chunk."""
    g = _g()
    n, nf = 4144 + 37, 26
    ip = g.TRMInputParameters(44100.0)
    whole = g.sweep_synthesize(ip, nf, n, seed=5, first_index=0, precision=g.TRM_PRECISION_FP32)
    assert whole["launches"] == 10                       # two chunks x (generator + 3 stages + checksum)
    parts = [g.sweep_synthesize(ip, nf, k1 - k0, seed=5, first_index=k0, precision=g.TRM_PRECISION_FP32)
             for k0, k1 in ((0, 1000), (1000, 4100), (4100, n))]
    assert np.array_equal(np.concatenate([p["checksums"] for p in parts]), whole["checksums"])
    assert np.array_equal(np.concatenate([p["maxima"] for p in parts]), whole["maxima"])
    assert len(set(whole["checksums"].tolist())) > n - 3          # different tracks, different audio


@pytest.mark.parametrize("precision", [0, 1])
def test_float32_frames_give_identical_results(precision):
    """TRM_FRAMES_F32: the same frames as rows of 16 floats (Monet's frames ARE floats, EventList.m:968-1002): half the
    upload, widened on the device, identical bytes out -- on the plain path and on the time-split path (>= 512 equal frame
    counts), ragged batches included."""
    g = _g()
    from gnuspeech_b200 import workloads as W, _native as N
    for n_frames in ([600] * 9, [37, 2, 300, 1, 64]):
        n = len(n_frames)
        frames = W.random_walk_ragged(n_frames, seed=3)
        assert np.array_equal(frames, frames.astype(np.float32).astype(np.float64))
        ip = g.TRMInputParameters(44100.0)
        out = []
        for fmt in (N.TRM_FRAMES_F64, N.TRM_FRAMES_F32):
            b = g.TRMBatch(ip, n_frames, precision=precision)
            b.set_frame_format(fmt)
            pcm = np.zeros(b.layout.total_pcm_samples, np.int16)
            smp = np.zeros(b.layout.total_out_samples, b.sample_dtype)
            b.synthesize(frames if fmt == N.TRM_FRAMES_F64 else np.ascontiguousarray(frames, np.float32), pcm_out=pcm, samples_out=smp, devices=[0])
            out.append((b.numberSamples.copy(), b.maximumSampleValues.copy(), pcm, smp, b))
        (ns0, mx0, p0, s0, b0), (ns1, mx1, p1, s1, _) = out
        assert np.array_equal(ns0, ns1) and np.array_equal(mx0, mx1)
        for u in range(n):
            o, k, c = b0.outOffsets[u], ns0[u], b0.pcmOffsets[u]
            assert np.array_equal(s0[o:o + k], s1[o:o + k]) and np.array_equal(p0[c:c + k], p1[c:c + k]), u


def test_copy_probe_reports_a_time():
    g = _g()
    from gnuspeech_b200 import _native as N
    a, b = g.PinnedArray(1 << 22, np.uint8), g.PinnedArray(1 << 23, np.uint8)
    ms = C.c_double(0.0)
    assert N.lib().TRMCopyProbe(0, a.ptr, 1 << 22, b.ptr, 1 << 23, 3, C.byref(ms)) == 0
    assert 0.0 < ms.value < 1000.0
    a.free(); b.free()
