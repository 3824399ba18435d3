"""world_size-2 `gloo` test of the multi-GPU plumbing (CPU): ranks take disjoint shards of the utterance stream,
generate their own frames from the shared counter-based generator, agree on the timing reduction bench.py uses
(MAX over ranks), and need no data-path collective."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import gnuspeech_b200 as g
    import oracle_lib as O
    from gnuspeech_b200 import sharding, workloads as W

    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    n_total, nf = 7, 12
    lo, hi = sharding.shard_range(n_total, rank, world)
    frames = W.random_walk(hi - lo, nf, seed=4, first_index=lo)
    # plan this rank's shard with the product's host library (no GPU needed to plan)
    b = g.TRMBatch(g.TRMInputParameters(44100.0), [nf] * (hi - lo))
    audio = torch.tensor([b.layout.audio_seconds], dtype=torch.float64)
    dist.all_reduce(audio, op=dist.ReduceOp.SUM)              # bench.py: whole-job audio seconds
    t = torch.tensor([0.5 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)                  # bench.py: time = max over ranks
    dist.barrier()
    # the checker runs the shard on the CPU so the union can be compared with a single-process run
    ns, mx, cs = O.synthesize_batch(O.male_voice(44100.0), frames, [nf] * (hi - lo), threads=2)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), lo=lo, hi=hi, frames=frames, ns=ns, mx=mx, cs=cs,
             audio=audio.numpy(), t=t.numpy())
    dist.destroy_process_group()


def test_two_rank_sharding(tmp_path):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    from gnuspeech_b200 import sharding, workloads as W

    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    n_total, nf = 7, 12
    whole = W.random_walk(n_total, nf, seed=4)
    ns, mx, cs = O.synthesize_batch(O.male_voice(44100.0), whole, [nf] * n_total, threads=2)
    covered = []
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        lo, hi = int(z["lo"]), int(z["hi"])
        covered += list(range(lo, hi))
        assert np.array_equal(z["frames"], whole[lo * nf:hi * nf])      # same tracks without exchanging data
        assert np.array_equal(z["ns"], ns[lo:hi]) and np.array_equal(z["mx"], mx[lo:hi]) and np.array_equal(z["cs"], cs[lo:hi])
        assert abs(float(z["audio"][0]) - n_total * (nf - 1) / 250.0) < 1e-12
        assert float(z["t"][0]) == 1.5
    assert covered == list(range(n_total))
    assert sharding.shard_bounds(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert sharding.weak_scaling_first_index(3, 4096) == 12288
