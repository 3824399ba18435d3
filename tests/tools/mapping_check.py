"""Both waveguide mappings (TRM_TUBE_MAPPING=sections / utterances) on the same batch: each against the oracle on a few
utterances, against each other on all of them, and the tube-kernel time of both.
usage: python tools/mapping_check.py [n_utt] [seconds] [n_check]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402
import gnuspeech_b200 as g  # noqa: E402
from gnuspeech_b200 import workloads as W  # noqa: E402
from gnuspeech_b200 import _native as N  # noqa: E402


def run(ip, frames, nfl, precision, mapping):
    os.environ["TRM_TUBE_MAPPING"] = mapping
    b = g.TRMBatch(ip, nfl, precision=precision)
    lay = b.layout
    pcm = np.zeros(max(1, lay.total_pcm_samples), np.int16)
    smp = np.zeros(max(1, lay.total_out_samples), b.sample_dtype)
    tube = np.zeros(max(1, b.tubeElements), b.sample_dtype)
    b.synthesize_debug(frames, pcm, smp, tube)
    return b, pcm, smp, tube


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    sec = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    n_check = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    ragged = len(sys.argv) > 4
    rng = np.random.default_rng(5)
    nfl = [int(sec * 250) + 1] * n if not ragged else [int(x) for x in rng.integers(2, int(sec * 250) + 2, n)]
    ip = g.TRMInputParameters(44100.0)
    frames = W.random_walk_ragged(nfl, seed=3) if ragged else W.random_walk(n, nfl[0], seed=3)
    off = np.concatenate(([0], np.cumsum(nfl)))
    for prec, name in ((g.TRM_PRECISION_FP64, "fp64"), (g.TRM_PRECISION_FP32, "fp32")):
        res = {}
        for mapping in ("sections", "utterances"):
            t0 = time.perf_counter()
            res[mapping] = run(ip, frames, nfl, prec, mapping)
            print("%s %-10s synth %.3f s" % (name, mapping, time.perf_counter() - t0), flush=True)
        bs, _, ss, ts = res["sections"]
        bu, pu, su, tu = res["utterances"]
        assert (bs.numberSamples == bu.numberSamples).all()
        pk = max(float(np.abs(ts).max()), 1e-300)
        print("%s  tube: sections vs utterances max diff %.3e of peak, nan %d / %d" % (
            name, float(np.nanmax(np.abs(ts.astype(np.float64) - tu.astype(np.float64)))) / pk, int(np.isnan(tu).sum()), int(np.isnan(ts).sum())))
        ns, oo, to = bu.numberSamples, bu.outOffsets, bu.tubeOffsets
        for u in list(range(min(n_check, n))) + ([n - 1] if n > n_check else []):
            ref = O.synthesize(ip, frames[off[u]:off[u + 1]], want_tube=True)
            if ref.numberSamples == 0:
                continue
            y = su[oo[u]:oo[u] + ns[u]].astype(np.float64)
            tt = tu[to[u]:to[u] + ref.tube.size].astype(np.float64)
            peak = ref.maximumSampleValue
            e = np.abs(y - ref.samples).max() / peak
            et = np.abs(tt - ref.tube).max() / np.abs(ref.tube).max()
            print("%s  utt %4d: n=%d out err %.3e tube err %.3e snr %.1f dB" % (name, u, ns[u], e, et, O.snr_db(ref.samples, y)))
        # kernel timing through the resident path
        import torch
        for mapping in ("sections", "utterances"):
            os.environ["TRM_TUBE_MAPPING"] = mapping
            b = g.TRMBatch(ip, nfl, precision=prec)
            pin = g.PinnedArray(frames.shape, np.float64)
            pin.array[:] = frames
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
            print("%s %-10s tube kernel %.3f ms" % (name, mapping, e0.elapsed_time(e1) / 3), flush=True)
            r.free()


if __name__ == "__main__":
    main()
