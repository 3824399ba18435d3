"""Timeline of the pipelined end-to-end loop of bench.py (3 calls in flight): every call prints its stage times on one
axis (TRM_TRACE=1, trm_cuda.cu).  usage: python tools/e2e_pipeline_trace.py [fp32] [steps]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gnuspeech_b200 as g
from gnuspeech_b200 import workloads as W
n, nf, depth = 4096, 2501, 3
prec = g.TRM_PRECISION_FP32 if "fp32" in sys.argv else g.TRM_PRECISION_FP64
steps = int([a for a in sys.argv[1:] if a.isdigit()][0]) if [a for a in sys.argv[1:] if a.isdigit()] else 5
ip = g.TRMInputParameters(44100.0)
frames = g.PinnedArray((n * nf, 16), np.float64)
W.random_walk(n, nf, seed=1, out=frames.array)
batches = [g.TRMBatch(ip, [nf] * n, precision=prec) for _ in range(depth)]
pcms = [g.PinnedArray(int(batches[0].layout.total_pcm_samples), np.int16) for _ in range(depth)]
for _ in range(2):
    tk = [batches[d].synthesize_async(frames, pcm_out=pcms[d], devices=[0]) for d in range(depth)]
    for t in tk:
        t.wait()
os.environ["TRM_TRACE"] = "1"
t0 = time.perf_counter()
tickets = []
for k in range(steps):
    if len(tickets) == depth:
        tickets.pop(0).wait()
    tickets.append(batches[k % depth].synthesize_async(frames, pcm_out=pcms[k % depth], devices=[0]))
    print("submitted %d at %.1f ms" % (k, 1e3 * (time.perf_counter() - t0)), file=sys.stderr, flush=True)
while tickets:
    tickets.pop(0).wait()
print("total %.1f ms, %.1f ms/step" % (1e3 * (time.perf_counter() - t0), 1e3 * (time.perf_counter() - t0) / steps), file=sys.stderr)
