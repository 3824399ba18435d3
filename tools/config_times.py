"""Throughput of the other BASELINE configs (parity-test cases, not bench lines): configs[2] full static grid (65,536 x
0.5 s), configs[3] 256 mixed-length utterances (5-60 s, 44.1 / 22.05 kHz), configs[4] slice (32,768 x 2 s) -- kernels
only (device-resident, CUDA events) and through the blocking host API."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gnuspeech_b200 as g
from gnuspeech_b200 import workloads as W, _native as N


def run(name, ips, frames, n_frames, prec):
    b = g.TRMBatch(ips, n_frames, precision=prec)
    audio = float(b.layout.audio_seconds)
    pin = g.PinnedArray(frames.shape, np.float64)
    pin.array[:] = frames
    r = b.make_resident(pin, device=0)
    st = torch.cuda.current_stream()
    for _ in range(2):
        r.run(st.cuda_stream)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record(st); r.run_stage(N.TRM_STAGE_TUBE, st.cuda_stream); ev[1].record(st)
    r.run_stage(N.TRM_STAGE_SRC, st.cuda_stream); ev[2].record(st); r.run_stage(N.TRM_STAGE_PCM, st.cuda_stream); ev[3].record(st)
    torch.cuda.synchronize()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]
    r.free()
    pcm = g.PinnedArray(int(b.layout.total_pcm_samples), np.int16)
    b.synthesize(pin, pcm_out=pcm, devices=[0])
    t0 = time.perf_counter()
    for _ in range(2):
        b.synthesize(pin, pcm_out=pcm, devices=[0])
    dt = (time.perf_counter() - t0) / 2
    print("%-44s %s: audio %8.0f s  kernels %7.2f + %6.2f + %5.2f ms = %9.0f audio-s/s   blocking e2e %7.1f ms = %8.0f audio-s/s" % (
        name, "fp64" if prec == 0 else "fp32", audio, ms[0], ms[1], ms[2], audio / (sum(ms) * 1e-3), dt * 1e3, audio / dt), flush=True)
    pcm.free(); pin.free()


ip = g.TRMInputParameters(44100.0)
for prec in (0, 1):
    n, nf = 65536, 126
    run("configs[2] 65,536-point grid x 0.5 s", ip, W.grid(range(n), nf), [nf] * n, prec)
    rng = np.random.default_rng(4)
    nfl = [int(x) for x in rng.integers(5 * 250, 60 * 250 + 1, 256)]
    ips = [g.TRMInputParameters(44100.0 if u % 2 == 0 else 22050.0) for u in range(256)]
    run("configs[3] 256 utterances 5-60 s mixed rates", ips, W.random_walk_ragged(nfl, seed=9), nfl, prec)
    n, nf = 32768, 501
    run("configs[4] slice: 32,768 x 2 s", ip, W.random_walk(n, nf, seed=5), [nf] * n, prec)
