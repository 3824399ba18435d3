"""Latency / throughput of streaming synthesis: n streams, pushes of m frames (4 ms each)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gnuspeech_b200 as g
from gnuspeech_b200 import workloads as W
ip = g.TRMInputParameters(44100.0)
for prec, name in ((0, "fp64"), (1, "fp32")):
    for n, m in ((1, 5), (64, 5), (1024, 10), (4096, 25), (4096, 250)):
        pushes = 12
        frames = W.random_walk(n, m * pushes + 1, seed=3).reshape(n, -1, 16)
        st = g.TRMStream(n, ip, precision=prec, max_frames_per_push=m + 1)
        st.push(frames[:, :m + 1])
        t = []
        for k in range(1, pushes):
            t0 = time.perf_counter()
            st.push(frames[:, 1 + k * m:1 + (k + 1) * m])
            t.append(time.perf_counter() - t0)
        st.free()
        dt = float(np.median(t))
        print("%s  %5d streams x %4d ms per push: %7.2f ms per push -> %9.0f audio-s/s (%.0fx real time per stream)" % (
            name, n, m * 4, dt * 1e3, n * m * 0.004 / dt, m * 0.004 / dt), flush=True)
