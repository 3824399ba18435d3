"""Times the waveguide stage alone (resident inputs). usage: wide_time.py n_utt seconds [precision]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gnuspeech_b200 as g
from gnuspeech_b200 import workloads as W, _native as N
n = int(sys.argv[1]); sec = float(sys.argv[2])
precs = sys.argv[3:] or ["fp64", "fp32"]
nf = int(sec * 250) + 1
ip = g.TRMInputParameters(44100.0)
pin = g.PinnedArray((n * nf, 16), np.float64)
W.random_walk(n, nf, seed=1, out=pin.array)
for p in precs:
    b = g.TRMBatch(ip, [nf] * n, precision=g.TRM_PRECISION_FP64 if p == "fp64" else g.TRM_PRECISION_FP32)
    r = b.make_resident(pin, device=0)
    st = torch.cuda.current_stream()
    for _ in range(2):
        r.run_stage(N.TRM_STAGE_TUBE, st.cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(3):
        r.run_stage(N.TRM_STAGE_TUBE, st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    print("%s mapping=%s debug=%s n=%d sec=%g tube %.3f ms" % (p, os.environ.get("TRM_TUBE_MAPPING"), os.environ.get("TRM_WIDE_DEBUG"), n, sec, e0.elapsed_time(e1) / 3), flush=True)
    r.free()
