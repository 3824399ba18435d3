import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gnuspeech_b200 as g
from gnuspeech_b200 import workloads as W
n, nf = 4096, 2501
prec = g.TRM_PRECISION_FP32 if "fp32" in sys.argv else g.TRM_PRECISION_FP64
ip = g.TRMInputParameters(44100.0)
frames = g.PinnedArray((n * nf, 16), np.float64)
W.random_walk(n, nf, seed=1, out=frames.array)
b = g.TRMBatch(ip, [nf] * n, precision=prec)
pcm = g.PinnedArray(int(b.layout.total_pcm_samples), np.int16)
os.environ.pop("TRM_TRACE", None)
b.synthesize(frames, pcm_out=pcm, devices=[0])
b.synthesize(frames, pcm_out=pcm, devices=[0])
os.environ["TRM_TRACE"] = "1"
b.synthesize(frames, pcm_out=pcm, devices=[0])
