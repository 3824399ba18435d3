import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gnuspeech_b200 as g
from gnuspeech_b200 import workloads as W

def run(ips, frames, nfl, prec):
    b = g.TRMBatch(ips, nfl, precision=prec)
    pcm = np.zeros(max(1, b.layout.total_pcm_samples), np.int16)
    b.synthesize(frames, pcm_out=pcm, devices=[0])
    return b, pcm

# some earlier, different work on lane 0 (other voices, other sizes)
ips = [g.TRMInputParameters(44100.0, length=15.0), g.TRMInputParameters(22050.0, length=10.0), g.TRMInputParameters(44100.0, waveform=1)]
run(ips, W.random_walk(3, 101, seed=33), [101] * 3, 0)
if len(sys.argv) > 1:
    run(g.TRMInputParameters(44100.0), W.random_walk(2048, 501, seed=5), [501] * 2048, 1)
n, nf = 96, 126
ip = g.TRMInputParameters(44100.0)
fa, fb = W.random_walk(n, nf, seed=41), W.random_walk(n, nf, seed=42)
ref = []
for fr in (fa, fb):
    b, pcm = run(ip, fr, [nf] * n, 0)
    ref.append((pcm.copy(), b.maximumSampleValues.copy(), b.pcmOffsets.copy(), b.numberSamples.copy()))
batches = [g.TRMBatch(ip, [nf] * n, precision=0) for _ in range(2)]
pcms = [np.zeros(batches[0].layout.total_pcm_samples, np.int16) for _ in range(2)]
for r in range(4):
    t = [batches[k].synthesize_async((fa, fb)[k], pcm_out=pcms[k], devices=[0]) for k in range(2)]
    for k in (1, 0):
        t[k].wait()
    for k in range(2):
        po, ns = ref[k][2], ref[k][3]
        bad = [u for u in range(n) if not np.array_equal(pcms[k][po[u]:po[u] + ns[u]], ref[k][0][po[u]:po[u] + ns[u]])]
        mbad = np.nonzero(batches[k].maximumSampleValues != ref[k][1])[0]
        print("round", r, "call", k, "pcm-bad utterances", bad[:10], len(bad), "max-bad", mbad[:10], len(mbad))
        if bad:
            u = bad[0]
            a, b_ = pcms[k][po[u]:po[u] + ns[u]].astype(int), ref[k][0][po[u]:po[u] + ns[u]].astype(int)
            d = np.nonzero(a != b_)[0]
            print("   first diffs at", d[:8], a[d[:8]], b_[d[:8]], "count", d.size)
        pcms[k][:] = 0
