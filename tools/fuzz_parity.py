"""Randomised parity sweep on the GPU: voices drawn over the reference GUI's ranges, tracks = random walks with
edge values patched in (parameters exactly at amplitude()'s clamps, closed velum, frication tap at the tube's ends,
pitch extremes, constant stretches).  Checks, per utterance:
  conformance vs strict   <= 1e-9 of peak on the output samples, PCM +-1 LSB, maxima to 1e-9
  strict vs CPU oracle    <= 1e-9 (in practice <= 1e-12) on a random subset (the oracle is one CPU thread)
  FP32 fast vs strict     SNR >= 60 dB (the fast mode's contract)
usage: python tools/fuzz_parity.py [n_utterances=400] [seed=1] [oracle_subset=24]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gnuspeech_b200 as g  # noqa: E402
import oracle_lib as O  # noqa: E402
from gnuspeech_b200 import workloads as W  # noqa: E402


def voice(rng):
    kw = dict(
        length=float(rng.uniform(10.0, 20.0)), temperature=float(rng.uniform(25.0, 40.0)),
        lossFactor=float(rng.uniform(0.0, 5.0)), apScale=float(rng.uniform(1.0, 5.0)),
        mouthCoef=float(rng.uniform(1000.0, 8000.0)), noseCoef=float(rng.uniform(1000.0, 8000.0)),
        throatCutoff=float(rng.uniform(500.0, 3000.0)), throatVol=float(rng.uniform(0.0, 24.0)),
        breathiness=float(rng.uniform(0.0, 10.0)), mixOffset=float(rng.uniform(30.0, 60.0)),
        tp=float(rng.uniform(20.0, 45.0)), tnMin=float(rng.uniform(10.0, 20.0)), tnMax=float(rng.uniform(25.0, 40.0)),
        waveform=int(rng.integers(0, 2)), usesModulation=int(rng.integers(0, 2)),
        channels=int(rng.integers(1, 3)), balance=float(rng.uniform(-1.0, 1.0)), volume=float(rng.uniform(30.0, 60.0)),
        noseRadius=[0.0] + [float(x) for x in rng.uniform(0.2, 3.0, 5)],
    )
    return g.TRMInputParameters(44100.0 if rng.random() < 0.7 else 22050.0, **kw)


def patch_edges(fr, rng):
    """fr: (nf, 16) frames of one utterance: 0 pitch, 1 glotVol, 2 aspVol, 3 fricVol, 4 fricPos, 5 fricCF, 6 fricBW,
    7..14 radii, 15 velum (TRMParameters order)."""
    nf = fr.shape[0]
    for _ in range(int(rng.integers(0, 4))):
        a = int(rng.integers(0, nf))
        b = min(nf, a + int(rng.integers(1, 12)))
        kind = int(rng.integers(0, 8))
        if kind == 0:
            fr[a:b, 1] = 60.0                    # glottal volume at the upper clamp
        elif kind == 1:
            fr[a:b, 1] = 0.0                     # ... and at the lower one
        elif kind == 2:
            fr[a:b, 3] = float(rng.choice([0.0, 60.0, 0.5]))
        elif kind == 3:
            fr[a:b, 4] = float(rng.choice([0.0, 7.0, 3.0, 6.999, 0.001]))
        elif kind == 4:
            fr[a:b, 15] = 0.0                    # closed velum
        elif kind == 5:
            fr[a:b, 0] = float(rng.choice([-12.0, 12.0, 0.0]))
        elif kind == 6:
            fr[a:b, 2] = float(rng.choice([0.0, 60.0]))
        else:
            fr[a:b, 7 + int(rng.integers(0, 8))] = 0.05   # a nearly closed section (not two adjacent: that is the NaN test)
    return fr


def run(n=400, seed=1, n_or=24):
    """Returns (worst, failures): worst[...] = (value, utterance) of each check."""
    rng = np.random.default_rng(seed)
    nfl = [int(x) for x in rng.integers(2, 400, n)]
    ips = [voice(rng) for _ in range(n)]
    frames = W.random_walk_ragged(nfl, seed=1000 + seed)
    off = np.concatenate(([0], np.cumsum(nfl)))
    for u in range(n):
        patch_edges(frames[off[u]:off[u + 1]], rng)
    res = {}
    for p in (g.TRM_PRECISION_FP64, g.TRM_PRECISION_FP64_STRICT, g.TRM_PRECISION_FP32):
        b = g.TRMBatch(ips, nfl, precision=p)
        smp = np.zeros(max(1, b.layout.total_out_samples), b.sample_dtype)
        tb = np.zeros(max(1, b.tubeElements), b.sample_dtype)
        pcm = np.zeros(max(1, b.layout.total_pcm_samples), np.int16)
        b.synthesize_debug(frames, pcm, smp, tb)
        res[p] = (b, smp, pcm)
    b0, s0, p0 = res[g.TRM_PRECISION_FP64]
    b2, s2, p2 = res[g.TRM_PRECISION_FP64_STRICT]
    b1, s1, _ = res[g.TRM_PRECISION_FP32]
    subset = set(int(x) for x in rng.choice(n, size=min(n_or, n), replace=False))
    worst = dict(cs=(0.0, -1), pcm=(0, -1), mx=(0.0, -1), so=(0.0, -1), snr=(1e9, -1))
    bad = 0
    for u in range(n):
        ns, o = int(b0.numberSamples[u]), int(b0.outOffsets[u])
        if ns == 0:
            continue
        y0, y2 = s0[o:o + ns], s2[o:o + ns]
        if not (np.isfinite(y2).all() and np.isfinite(y0).all()):
            if not np.array_equal(np.isnan(y0), np.isnan(y2)):
                print("utterance %d: NaN patterns differ" % u)
                bad += 1
            continue
        pk = max(float(np.abs(y2).max()), 1e-300)
        cs = float(np.abs(y0 - y2).max()) / pk
        worst["cs"] = max(worst["cs"], (cs, u))
        ch = 2 if ips[u].channels == 2 else 1
        po = int(b0.pcmOffsets[u])
        dp = int(np.abs(p0[po:po + ns * ch].astype(np.int32) - p2[po:po + ns * ch].astype(np.int32)).max())
        worst["pcm"] = max(worst["pcm"], (dp, u))
        m0, m2 = float(b0.maximumSampleValues[u]), float(b2.maximumSampleValues[u])
        worst["mx"] = max(worst["mx"], (abs(m0 - m2) / max(m2, 1e-300), u))
        snr = O.snr_db(y2, s1[int(b1.outOffsets[u]):int(b1.outOffsets[u]) + ns].astype(np.float64))
        worst["snr"] = min(worst["snr"], (snr, u))
        if u in subset:
            ref = O.synthesize(ips[u], frames[off[u]:off[u + 1]], want_tube=False)
            so = float(np.abs(y2 - ref.samples).max()) / max(ref.maximumSampleValue, 1e-300)
            worst["so"] = max(worst["so"], (so, u))
        if cs > 1e-9 or dp > 1:
            bad += 1
            print("utterance %d: conformance vs strict %.3e, PCM %d LSB (voice: length %.2f, rate %.0f, waveform %d)" %
                  (u, cs, dp, ips[u].length, ips[u].outputRate, ips[u].waveform))
    print("fuzz seed %d, %d utterances (%d frames): conformance-vs-strict %.2e (utt %d) | PCM %d LSB | maxima %.2e | "
          "strict-vs-oracle %.2e on %d | FP32 SNR >= %.1f dB (utt %d) | failures %d" %
          (seed, n, int(off[-1]), worst["cs"][0], worst["cs"][1], worst["pcm"][0], worst["mx"][0], worst["so"][0], len(subset),
           worst["snr"][0], worst["snr"][1], bad), flush=True)
    return worst, bad


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    n_or = int(sys.argv[3]) if len(sys.argv) > 3 else 24
    worst, bad = run(n, seed, n_or)
    if bad or worst["so"][0] > 1e-9 or worst["snr"][0] < 60.0:
        raise SystemExit(1)


if __name__ == "__main__":
    main()
