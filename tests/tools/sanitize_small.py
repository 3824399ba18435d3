"""Small end-to-end runs of every kernel for compute-sanitizer (memcheck / racecheck): both waveguide mappings, ragged
batch, up- and down-sampling converter, PCM, frame generator."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
import gnuspeech_b200 as g
from gnuspeech_b200 import workloads as W
n_frames = [26, 1, 40, 2, 33, 51, 17]
frames = W.random_walk_ragged(n_frames, seed=3)
ips = [g.TRMInputParameters(44100.0), g.TRMInputParameters(22050.0, length=15.0), g.TRMInputParameters(22050.0, length=10.0),
       g.TRMInputParameters(44100.0), g.TRMInputParameters(22050.0, channels=2, balance=0.3), g.TRMInputParameters(44100.0, waveform=1),
       g.TRMInputParameters(44100.0, length=7.5)]
for mapping in ("sections", "utterances"):
    os.environ["TRM_TUBE_MAPPING"] = mapping
    for prec in (0, 1):
        b = g.TRMBatch(ips, n_frames, precision=prec)
        pcm = np.zeros(b.layout.total_pcm_samples, np.int16)
        smp = np.zeros(b.layout.total_out_samples, b.sample_dtype)
        b.synthesize(frames, pcm_out=pcm, samples_out=smp, devices=[0])
        print(mapping, prec, "ok", int(np.abs(pcm).max()))
evs = [O.synthetic_event_list(10 + u, 0.2) for u in range(5)]
nf = [g.event_list_frame_count(e) for e in evs]
b = g.TRMBatch(g.TRMInputParameters(44100.0), nf, precision=1)
pcm = np.zeros(b.layout.total_pcm_samples, np.int16)
b.synthesize_events(np.concatenate(evs), [len(e) for e in evs], g.TRMFrameGeneration(), pcm_out=pcm)
print("events ok", int(np.abs(pcm).max()))
