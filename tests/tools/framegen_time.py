"""Times the control-frame generator and the events -> PCM path on configs[1]-sized input (4096 x 10 s)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
import gnuspeech_b200 as g
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sec = float(sys.argv[2]) if len(sys.argv) > 2 else 10.0
base = [O.synthetic_event_list(500 + u, sec) for u in range(64)]
evs = [base[u % 64] for u in range(n)]
counts = [len(e) for e in evs]
allev = np.concatenate(evs)
n_frames = [g.event_list_frame_count(e) for e in base]
n_frames = [n_frames[u % 64] for u in range(n)]
ip = g.TRMInputParameters(44100.0)
fg = g.TRMFrameGeneration()
for prec, name in ((g.TRM_PRECISION_FP64, "fp64"), (g.TRM_PRECISION_FP32, "fp32")):
    b = g.TRMBatch(ip, n_frames, precision=prec)
    pcm = g.PinnedArray(int(b.layout.total_pcm_samples), np.int16)
    audio = float(b.layout.audio_seconds)
    print("events %.1f MB, frames %.1f MB, audio %.0f s" % (allev.nbytes / 1e6, b.layout.total_frames * 128 / 1e6, audio))
    for _ in range(2):
        b.synthesize_events(allev, counts, fg, pcm_out=pcm)
    t0 = time.perf_counter()
    for _ in range(3):
        b.synthesize_events(allev, counts, fg, pcm_out=pcm)
    dt = (time.perf_counter() - t0) / 3
    print("%s events -> PCM (blocking): %.1f ms per call, %.0f audio-s/s" % (name, dt * 1e3, audio / dt))
    t0 = time.perf_counter()
    for _ in range(3):
        b.generate_frames(allev, counts, fg)
    print("%s generate_frames incl. D2H of frames: %.1f ms" % (name, (time.perf_counter() - t0) / 3 * 1e3))
    pcm.free()
