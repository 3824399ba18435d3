"""Copy-only ceiling of the end-to-end path: every rank copies one bench step's bytes -- frames up (float32 rows, or
double rows with --f64), PCM16 down -- at the same time on two streams, with no kernels (TRMCopyProbe, include/trm.h), all
ranks concurrently.  Prints one JSON line (rank 0): per-rank and aggregate GB/s and the audio-s/s the end-to-end number
of configs[1] cannot exceed on this host.

    python tools/pcie_ceiling.py                                        # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_ceiling.py
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import gnuspeech_b200 as g  # noqa: E402
from gnuspeech_b200 import _native as N  # noqa: E402

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_utt, nf = 4096, 2501
ip = g.TRMInputParameters(44100.0)
lay = g.TRMBatch(ip, [nf] * n_utt).layout
h2d = int(lay.total_frames) * (128 if "--f64" in sys.argv else 64)
d2h = int(lay.out_samples) * 2
a, b = g.PinnedArray(h2d, np.uint8), g.PinnedArray(d2h, np.uint8)
a.array[:] = 1
res = {}
for name, up, down in (("both", h2d, d2h), ("h2d_only", h2d, 0), ("d2h_only", 0, d2h)):
    ms = C.c_double(0.0)
    if dist is not None:
        dist.barrier()
    N.check(N.lib().TRMCopyProbe(local, a.ptr, up, b.ptr, down, 5, C.byref(ms)), "TRMCopyProbe")
    t = torch.tensor([ms.value], dtype=torch.float64, device="cuda")
    if dist is not None:
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        per = [float(x.item()) for x in allt]
    else:
        per = [ms.value]
    worst = max(per)
    res[name] = {"ms_per_step_max_over_ranks": worst, "ms_per_rank": per,
                 "aggregate_gbs": world * (up + down) / (worst * 1e-3) / 1e9,
                 "audio_s_per_s_ceiling": world * float(lay.audio_seconds) / (worst * 1e-3)}
if rank == 0:
    print(json.dumps({"n_gpus": world, "h2d_bytes_per_rank": h2d, "d2h_bytes_per_rank": d2h, "workload": "configs[1] byte counts per rank",
                      "host_cpus": len(os.sched_getaffinity(0)), **res}))
if dist is not None:
    dist.destroy_process_group()
