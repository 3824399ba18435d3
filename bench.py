#!/usr/bin/env python
"""bench.py -- TRM synthesis throughput on B200 (BASELINE.json metric: synthesized audio-seconds per
wall-second, batched).

Workload (config.workload): BASELINE.json configs[1] per GPU -- 4096 synthetic random-walk utterances x 10 s
(2501 control frames each, male voice, 44.1 kHz mono), weak scaling: every rank synthesizes its own 4096.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision fp64|fp32] [--impl reference]

One "step" = one pass of the hot path (waveguide -> resampler -> PCM) over the whole batch.
  value     : frames resident in HBM, the three kernels launched back to back on torch's current stream,
              timed with CUDA events on that stream (max over ranks).
  e2e       : same batch through the public C API TRMBatchSynthesize with HOST buffers (pinned): H2D of the
              frames and D2H of the PCM inside the timed region.
  roofline  : dominant kernel (waveguide) -- FP-pipe bound, so the fraction is achieved FLOP/s over an FMA
              peak MEASURED live on this device (MEASURED_PEAKS.json has no CUDA-core number); the resampler
              is reported against the measured HBM peak in roofline_src.
  cpu_baseline : the CPU oracle (a C restatement of the reference, kind "port"), one utterance per thread on all
              host cores, on a bounded sample of the same workload.
--impl reference runs only that CPU arm, sized to finish in minutes, and prints its own JSON line.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# the host pipeline keeps many streams in flight; must be set before CUDA is initialised (by torch, below)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import numpy as np  # noqa: E402

METRIC = "synthesized audio-sec per wall-sec"
UNIT = "audio-s/s"
FLOP_PER_TUBE_SAMPLE = 390.0     # algorithmic flops per tube-rate sample
FLOP_PER_OUT_SAMPLE = 110.0      # up-sampling converter, per output sample
# DRAM traffic per unit measured with `ncu --set full` (dram__bytes_read.sum + dram__bytes_write.sum of the
# 4096 x 1 s launch, profiles/prof_r1b_{fp64,fp32}_summary.txt) -- linear in the number of samples:
#   waveguide: bytes per tube-rate sample, resampler / PCM: bytes per output sample
TRAFFIC_PER_UNIT = {"fp64": {"tube": 8.95, "src": 12.08, "pcm": 9.96}, "fp32": {"tube": 5.04, "src": 5.85, "pcm": 5.85}}


def bind_to_gpu_numa_node(torch, local_rank, world):
    """Several ranks share the host: keep this rank's threads (and with them the pinned buffers it allocates and the
    staging copies libtrm makes) on the CPUs next to its GPU, so that every GPU's PCIe traffic stays on its own socket.
    Only under torchrun; silently skipped when sysfs does not tell."""
    if world <= 1 or os.environ.get("TRM_NO_NUMA_BIND"):
        return
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/local_cpulist" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        cpus = set()
        for part in open(path).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            print("[bench] rank %d: bound to %d CPUs local to GPU %d" % (local_rank, len(cpus), local_rank), file=sys.stderr)
    except (OSError, ValueError, AttributeError):
        pass


def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(object):
    """Samples SM clocks and throttle reasons with nvidia-smi while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def oracle_throughput(ip, n_frames, n_utt, threads, seed, first_index):
    """Times the CPU oracle (one utterance per thread) on n_utt utterances of the workload; returns
    (audio_s_per_s, seconds)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    from gnuspeech_b200 import workloads as W
    frames = W.random_walk(n_utt, n_frames, seed=seed, first_index=first_index)
    t0 = time.perf_counter()
    O.synthesize_batch(ip, frames, [n_frames] * n_utt, flags=0, threads=threads)
    dt = time.perf_counter() - t0
    audio = n_utt * (n_frames - 1) / float(ip.controlRate)
    return audio / dt, dt


def run_reference_arm(args):
    """--impl reference: the reference's CPU path (oracle port of Frameworks/Tube; the Objective-C original
    cannot be built here and TRAcT/tube.c is one utterance per process), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import gnuspeech_b200 as g
    from gnuspeech_b200 import build as B
    B.build_oracle()
    ip = g.TRMInputParameters(44100.0)
    n_frames = int(args.seconds * 250) + 1
    cores = host_cores()
    # bounded sample per step: 16 utterances per core of the same 10 s random-walk workload (~1 s wall, ~15-20 s CPU)
    n_utt = max(cores * 16, 32) if args.sample_utterances <= 0 else args.sample_utterances
    for _ in range(args.warmup):
        oracle_throughput(ip, n_frames, min(n_utt, cores), cores, args.seed, 0)
    t_total, audio_total = 0.0, 0.0
    for k in range(args.steps):
        v, dt = oracle_throughput(ip, n_frames, n_utt, cores, args.seed, k * n_utt)
        t_total += dt
        audio_total += v * dt
    value = audio_total / t_total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "configs[1] sample: %d random-walk utterances x %g s per step (of 4096 x 10 s), male voice, "
                               "250 Hz control frames, 44.1 kHz mono" % (n_utt, args.seconds)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d utterances x %g s per step, one utterance per thread, reference-faithful "
                                   "per-sample wavetable rewrite" % (n_utt, args.seconds)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="fp64", choices=["fp64", "fp32"])
    ap.add_argument("--utterances", type=int, default=4096, help="utterances per GPU")
    ap.add_argument("--seconds", type=float, default=10.0, help="seconds of audio per utterance")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--sample-utterances", type=int, default=0, help="CPU arm: utterances per step (0 = 16 per core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--also-fp32", action="store_true", help="add a fast_mode object measured the same way in FP32")
    args = ap.parse_args()

    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import gnuspeech_b200 as g
    from gnuspeech_b200 import _native as N
    from gnuspeech_b200 import workloads as W

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the TRM path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    bind_to_gpu_numa_node(torch, local_rank, world)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when the communicator is created; stdout carries the JSON line only
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            warm = torch.zeros(1, device="cuda")
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    n_utt = args.utterances
    n_frames = int(args.seconds * 250) + 1
    ip = g.TRMInputParameters(44100.0)
    cores = host_cores()

    # ---- inputs: this rank's utterances, in pinned host memory --------------------------------------------
    frames = g.PinnedArray((n_utt * n_frames, 16), np.float64)
    W.random_walk(n_utt, n_frames, seed=args.seed, first_index=rank * n_utt, out=frames.array)

    def measure(precision):
        prec = g.TRM_PRECISION_FP64 if precision == "fp64" else g.TRM_PRECISION_FP32
        batch = g.TRMBatch(ip, [n_frames] * n_utt, precision=prec)
        lay = batch.layout
        audio_s = float(lay.audio_seconds)
        esz = 8 if precision == "fp64" else 4
        stream = torch.cuda.current_stream()
        sh = stream.cuda_stream

        # ---- value: HBM-resident inputs, kernels only --------------------------------------------------------
        res = batch.make_resident(frames, device=local_rank)
        for _ in range(args.warmup):
            res.run(sh)
        torch.cuda.synchronize()
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
        sampler = ClockSampler(local_rank)
        barrier()
        torch.cuda.synchronize()
        sampler.start()
        for k in range(args.steps):
            ev[k][0].record(stream)
            res.run_stage(N.TRM_STAGE_TUBE, sh)
            ev[k][1].record(stream)
            res.run_stage(N.TRM_STAGE_SRC, sh)
            ev[k][2].record(stream)
            res.run_stage(N.TRM_STAGE_PCM, sh)
            ev[k][3].record(stream)
        torch.cuda.synchronize()
        barrier()
        clocks = sampler.stop()
        total_ms = ev[0][0].elapsed_time(ev[-1][3])
        stage_ms = [float(np.mean([ev[k][i].elapsed_time(ev[k][i + 1]) for k in range(args.steps)])) for i in range(3)]
        total_ms = max_over_ranks(total_ms)
        ms_per_step = total_ms / args.steps
        audio_all = sum_over_ranks(audio_s)
        value = audio_all / (ms_per_step * 1e-3)
        # keep a result for the sanity check below
        maxima = np.zeros(n_utt, np.float64)
        res.fetch(None, None, maxima, None)
        res.free()

        # ---- e2e: public API, host buffers, copies inside the timed region ----------------------------------
        # Every step = one TRMBatch call on this step's pinned host frames -> this step's pinned host PCM16.  A caller
        # with a stream of batches keeps two calls in flight (TRMBatchSynthesizeAsync / TRMBatchWait, include/trm.h):
        # the PCM copy-out of step k overlaps the kernels of step k+1.  All K steps' H2D, kernels and D2H complete
        # inside the timed region (the clock stops after the last TRMBatchWait).  The strictly serial form (one
        # blocking TRMBatchSynthesize per step) is measured too and reported as e2e.blocking.
        e2e = None
        if not args.no_e2e:
            depth = 3
            batches = [batch] + [g.TRMBatch(ip, [n_frames] * n_utt, precision=prec) for _ in range(depth - 1)]
            pcms = [g.PinnedArray(int(lay.total_pcm_samples), np.int16) for _ in range(depth)]
            for _ in range(min(args.warmup, 2)):         # warms both context lanes (arenas, pinned staging)
                tk = [batches[d].synthesize_async(frames, pcm_out=pcms[d], devices=[local_rank]) for d in range(depth)]
                for t in tk:
                    t.wait()
            torch.cuda.synchronize()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                batches[0].synthesize(frames, pcm_out=pcms[0], devices=[local_rank])
            torch.cuda.synchronize()
            dt_block = time.perf_counter() - t0
            barrier()
            dt_block = max_over_ranks(dt_block)
            assert np.array_equal(batches[0].maximumSampleValues, maxima), "e2e and resident paths disagree"
            t0 = time.perf_counter()
            tickets = []
            for k in range(args.steps):
                if len(tickets) == depth:
                    tickets.pop(0).wait()
                tickets.append(batches[k % depth].synthesize_async(frames, pcm_out=pcms[k % depth], devices=[local_rank]))
            while tickets:
                tickets.pop(0).wait()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            barrier()
            dt = max_over_ranks(dt)
            e2e = {"value": audio_all / (dt / args.steps), "unit": UNIT,
                   "h2d_bytes_per_step": int(lay.total_frames) * 128 * world,
                   "d2h_bytes_per_step": int(lay.out_samples) * 2 * world,
                   "ms_per_step": 1e3 * dt / args.steps,
                   "gpu_launches_per_step": int(batches[0].kernelLaunches),
                   "api": "TRMBatchSynthesizeAsync / TRMBatchWait (include/trm.h), %d calls in flight, pinned host frames in, "
                          "pinned host PCM16 out; every step's copies and kernels finish inside the timed region" % depth,
                   "blocking": {"value": audio_all / (dt_block / args.steps), "ms_per_step": 1e3 * dt_block / args.steps,
                                "api": "TRMBatchSynthesize, one blocking call per step"}}
            for d in range(depth):
                assert np.array_equal(batches[d].maximumSampleValues, maxima), "e2e and resident paths disagree"
                assert int(np.abs(pcms[d].array[:1000].astype(np.int32)).max()) > 0
            po, ns = batches[0].pcmOffsets, batches[0].numberSamples
            for u in (0, n_utt // 2, n_utt - 1):
                assert np.array_equal(pcms[0].array[po[u]:po[u] + ns[u]], pcms[1].array[po[u]:po[u] + ns[u]]), "pipelined calls disagree"
            for pc in pcms:
                pc.free()
        return dict(value=value, ms_per_step=ms_per_step, stage_ms=stage_ms, clocks=clocks, e2e=e2e, lay=lay,
                    audio_all=audio_all, esz=esz, maxima=maxima)

    r = measure(args.precision)
    lay = r["lay"]
    hbm_peak, hbm_src = measured_peaks()

    # ---- roofline of the dominant kernel (waveguide): FP-pipe bound -> measured FMA peak on this device --------
    L = C.CDLL(N.LIBTRM_CUDA_PATH)
    L.trm_cuda_fp_peak.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
    peak_tf = C.c_double(0.0)
    L.trm_cuda_fp_peak(local_rank, 0 if args.precision == "fp64" else 1, 3, C.byref(peak_tf))
    tube_ms, src_ms, pcm_ms = r["stage_ms"]
    tube_flops = FLOP_PER_TUBE_SAMPLE * float(lay.tube_samples)
    tube_tf = tube_flops / (tube_ms * 1e-3) / 1e12
    esz = r["esz"]
    tube_bytes = float(lay.total_frames) * 128 + float(lay.tube_samples) * esz
    src_bytes = float(lay.tube_samples) * esz + float(lay.out_samples) * esz
    pcm_bytes = float(lay.out_samples) * (esz + 2)
    roofline = {
        "kernel": "tube_wide_kernel<%s>" % ("double" if args.precision == "fp64" else "float"),
        "bound": "fp64-pipe" if args.precision == "fp64" else "fp32-pipe",
        "achieved": tube_tf, "peak": peak_tf.value, "unit": "TFLOP/s", "frac": tube_tf / peak_tf.value if peak_tf.value else None,
        "peak_source": "FMA chain measured live on this device (trm_cuda_fp_peak); MEASURED_PEAKS.json has no CUDA-core peak",
        "flop_per_tube_sample": FLOP_PER_TUBE_SAMPLE, "ms_per_launch": tube_ms,
        "share_of_step": tube_ms / (tube_ms + src_ms + pcm_ms),
        "hbm_achieved_gbs": tube_bytes / (tube_ms * 1e-3) / 1e9,
        "algorithmic_bytes": tube_bytes, "traffic": TRAFFIC_PER_UNIT[args.precision]["tube"] * float(lay.tube_samples),
        "traffic_source": "ncu --set full dram bytes of the 4096 x 1 s launch, scaled by samples (profiles/)",
    }
    roofline_src = {
        "kernel": "src_kernel", "bound": "hbm", "achieved": src_bytes / (src_ms * 1e-3) / 1e9, "peak": hbm_peak,
        "unit": "GB/s", "frac": src_bytes / (src_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src,
        "ms_per_launch": src_ms, "share_of_step": src_ms / (tube_ms + src_ms + pcm_ms),
        "algorithmic_bytes": src_bytes, "traffic": TRAFFIC_PER_UNIT[args.precision]["src"] * float(lay.out_samples),
        "flops_tf": FLOP_PER_OUT_SAMPLE * float(lay.out_samples) / (src_ms * 1e-3) / 1e12,
    }
    roofline_pcm = {
        "kernel": "pcm_kernel", "bound": "hbm", "achieved": pcm_bytes / (pcm_ms * 1e-3) / 1e9, "peak": hbm_peak,
        "unit": "GB/s", "frac": pcm_bytes / (pcm_ms * 1e-3) / 1e9 / hbm_peak, "ms_per_launch": pcm_ms,
        "share_of_step": pcm_ms / (tube_ms + src_ms + pcm_ms),
        "algorithmic_bytes": pcm_bytes, "traffic": TRAFFIC_PER_UNIT[args.precision]["pcm"] * float(lay.out_samples),
    }

    fast = None
    if args.also_fp32 and args.precision == "fp64":
        f = measure("fp32")
        fast = {"dtype": "f32 (mixed: f64 pitch/phase, integer noise)", "value": f["value"], "ms_per_step": f["ms_per_step"],
                "stage_ms": f["stage_ms"], "e2e": f["e2e"]}

    # ---- CPU baseline on a bounded sample (rank 0, N=1 only) ------------------------------------------------
    cpu = None
    if not args.no_cpu_baseline and rank == 0 and world == 1:
        from gnuspeech_b200 import build as B
        B.build_oracle()
        n_s = min(n_utt, max(16 * cores, 32))           # ~15-20 s of CPU work
        v, dt = oracle_throughput(ip, n_frames, n_s, cores, args.seed, 0)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "seconds": dt,
               "sample": "%d of the %d utterances x %g s, one utterance per thread, reference-faithful per-sample "
                         "wavetable rewrite" % (n_s, n_utt, args.seconds)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64" if args.precision == "fp64" else "f32",
            "data": "synthetic",
            "config": {"workload": "configs[1]: %d random-walk utterances x %g s per GPU (%d control frames each, male voice, "
                                   "250 Hz control rate, 44.1 kHz mono PCM16)" % (n_utt, args.seconds, n_frames),
                       "utterances_per_gpu": n_utt, "audio_seconds_per_gpu": float(lay.audio_seconds),
                       "tube_samples_per_gpu": int(lay.tube_samples), "out_samples_per_gpu": int(lay.out_samples),
                       "precision_mode": args.precision,
                       "l2": "inputs_exceed_l2 (frames %.2f GB + tube-rate %.2f GB per step >> 126 MB)" % (
                           lay.total_frames * 128 / 1e9, lay.tube_samples * esz / 1e9),
                       "parallelism": "utterances sharded across GPUs, no collectives"},
            "clocks": r["clocks"],
            "e2e": r["e2e"],
            "gpu_launches": 3 * args.steps,
            "stage_ms": {"tube": tube_ms, "src": src_ms, "pcm": pcm_ms},
            "roofline": roofline, "roofline_src": roofline_src, "roofline_pcm": roofline_pcm,
            "cpu_baseline": cpu,
        }
        if fast is not None:
            line["fast_mode"] = fast
        print(json.dumps(line))
    frames.free()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
