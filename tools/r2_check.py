"""Round-2 quick check: FP64 conformance mode against the strict mode and the oracle, plus waveguide timings.
usage: r2_check.py [n_utt seconds]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import gnuspeech_b200 as g
import oracle_lib as O
from gnuspeech_b200 import workloads as W, _native as N

ip = g.TRMInputParameters(44100.0)


def run(ips, frames, nfl, prec, tube=True):
    b = g.TRMBatch(ips, nfl, precision=prec)
    smp = np.zeros(max(1, b.layout.total_out_samples), b.sample_dtype)
    tb = np.zeros(max(1, b.tubeElements), b.sample_dtype)
    pcm = np.zeros(max(1, b.layout.total_pcm_samples), np.int16)
    b.synthesize_debug(frames, pcm, smp, tb)
    return b, smp, tb, pcm


def compare(name, ips, frames, nfl, oracle_idx=()):
    res = {p: run(ips, frames, nfl, p) for p in (0, 2, 1)}
    b0, s0, t0, _ = res[0]
    b2, s2, t2, _ = res[2]
    b1, s1, t1, _ = res[1]
    off = np.concatenate(([0], np.cumsum(nfl)))
    worst_cs, worst_co, worst_so, snr32 = 0.0, 0.0, 0.0, 1e9
    for u in range(len(nfl)):
        n, o = b0.numberSamples[u], b0.outOffsets[u]
        if n == 0:
            continue
        pk = max(np.abs(s2[o:o + n]).max(), 1e-300)
        worst_cs = max(worst_cs, np.abs(s0[o:o + n] - s2[o:o + n]).max() / pk)
        if u in oracle_idx:
            ipu = ips[u] if isinstance(ips, (list, tuple)) else ips
            ref = O.synthesize(ipu, frames[off[u]:off[u + 1]], want_tube=False)
            worst_co = max(worst_co, np.abs(s0[o:o + n] - ref.samples).max() / ref.maximumSampleValue)
            worst_so = max(worst_so, np.abs(s2[o:o + n] - ref.samples).max() / ref.maximumSampleValue)
            snr32 = min(snr32, O.snr_db(ref.samples, s1[o:o + n].astype(np.float64)))
    print("%-34s conf-vs-strict %.2e | conf-vs-oracle %.2e | strict-vs-oracle %.2e | fp32 SNR %.1f dB" % (name, worst_cs, worst_co, worst_so, snr32), flush=True)


if len(sys.argv) <= 1 or sys.argv[1] == "check":
    nf = 251
    compare("static vowels", ip, np.concatenate([W.static_vowel(nf, 0), W.static_vowel(nf, 1)]), [nf, nf], (0, 1))
    n, nf = 12, 501
    compare("random walk 12 x 2 s", ip, W.random_walk(n, nf, seed=2), [nf] * n, range(12))
    rng = np.random.default_rng(11)
    nfl = [int(x) for x in rng.integers(2, 260, 75)] + [1, 2, 301]
    voices = [dict(), dict(length=15.0), dict(length=10.0, temperature=32.0), dict(waveform=1), dict(usesModulation=0, lossFactor=1.5)]
    ips = [g.TRMInputParameters(44100.0 if u % 3 else 22050.0, **voices[u % len(voices)]) for u in range(len(nfl))]
    compare("ragged mixed voices 78", ips, W.random_walk_ragged(nfl, seed=77), nfl, (0, 5, 33, 77))
    nf = 7501
    compare("30 s random walk", ip, W.random_walk(1, nf, seed=44), [nf], (0,))
    compare("30 s static vowel", ip, W.static_vowel(nf, 1), [nf], (0,))

if len(sys.argv) > 1 and sys.argv[1] != "check" or len(sys.argv) <= 1:
    n = int(sys.argv[1]) if len(sys.argv) > 2 else 4096
    sec = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
    nf = int(sec * 250) + 1
    pin = g.PinnedArray((n * nf, 16), np.float64)
    W.random_walk(n, nf, seed=1, out=pin.array)
    for p, name in ((0, "fp64"), (2, "fp64-strict"), (1, "fp32")):
        b = g.TRMBatch(ip, [nf] * n, precision=p)
        r = b.make_resident(pin, device=0)
        st = torch.cuda.current_stream()
        for _ in range(2):
            r.run(st.cuda_stream)
        ms = []
        for stage in (N.TRM_STAGE_TUBE, N.TRM_STAGE_SRC, N.TRM_STAGE_PCM):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(3):
                r.run_stage(stage, st.cuda_stream)
            e1.record(st)
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1) / 3)
        print("%-12s n=%d sec=%g: tube %.3f ms  src %.3f ms  pcm %.3f ms" % (name, n, sec, ms[0], ms[1], ms[2]), flush=True)
        r.free()
