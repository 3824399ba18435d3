"""Diagnostic run of one utterance on the GPU against the oracle (prints where the two diverge)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402
import gnuspeech_b200 as g  # noqa: E402
from gnuspeech_b200 import workloads as W  # noqa: E402


def report(name, a, b):
    peak = np.abs(b).max() if b.size else 0.0
    if a.shape != b.shape:
        print("%s: SHAPE %s vs %s" % (name, a.shape, b.shape))
        n = min(a.size, b.size)
        a, b = a[:n], b[:n]
    err = np.abs(a - b)
    bad = np.nonzero(~(err <= 1e-9 * peak))[0]
    print("%s: n=%d peak=%.6e max_err/peak=%.3e snr=%.1f dB first_bad=%s nan=%d" % (
        name, a.size, peak, (np.nanmax(err) / peak) if peak else 0.0, O.snr_db(b, a), bad[:5].tolist(), int(np.isnan(a).sum())))
    if bad.size:
        i = int(bad[0])
        lo = max(0, i - 2)
        print("   gpu:", a[lo:lo + 6])
        print("   ref:", b[lo:lo + 6])


def one(ip, frames, precision, tag):
    ref = O.synthesize(ip, frames)
    dl = g.TRMDataList()
    dl.setInputParameters(ip)
    dl.addParameters(frames)
    m = g.TRMTubeModel(dl, precision=precision)
    m.synthesize()
    print("== %s precision=%d numberSamples gpu=%d ref=%d max gpu=%.17g ref=%.17g" % (
        tag, precision, m.numberSamples, ref.numberSamples, m.maximumSampleValue, ref.maximumSampleValue))
    report("  tube", m.tubeSignal, ref.tube)
    report("  out ", m.resampledData, ref.samples)
    pcm_ref = O.pcm16(ip, ref.samples, ref.maximumSampleValue).astype(np.int32)
    pcm = m.pcm16().astype(np.int32)
    n = min(pcm.size, pcm_ref.size)
    d = np.abs(pcm[:n] - pcm_ref[:n])
    print("  pcm : differ=%d (>1: %d) max=%d" % (int((d > 0).sum()), int((d > 1).sum()), int(d.max()) if n else 0))


if __name__ == "__main__":
    ip = g.TRMInputParameters(44100.0)
    for prec in (g.TRM_PRECISION_FP64, g.TRM_PRECISION_FP32):
        one(ip, W.static_vowel(251, 1), prec, "static aa 1s")
        one(ip, W.random_walk(1, 501, seed=2), prec, "random walk 2s")
    ip22 = g.TRMInputParameters(22050.0)
    one(ip22, W.static_vowel(26, 0), g.TRM_PRECISION_FP64, "static a 0.1s @22050")
    # short tube -> down-sampling converter
    ipd = g.TRMInputParameters(22050.0, length=10.0)
    one(ipd, W.random_walk(1, 101, seed=3), g.TRM_PRECISION_FP64, "10 cm tube, down-sampling")
