"""N > 1 host logic on CPU: world_size-2 gloo.  Each rank channelizes the same tuner buffer (with the oracle --
no GPU here), keeps its slice of the bins, and the gathered slices must equal the unsharded result (level 2 of
SURVEY.md section 8e); stream / channel assignments must partition; the step time is the max over ranks."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
import siggen as sg
from sdrtrunk_b200 import sharding


def test_balanced_slices_partition():
    for n in (1, 7, 8, 400, 800, 1024):
        for world in (1, 2, 3, 4, 8):
            slices = [sharding.balanced_slice(n, r, world) for r in range(world)]
            assert slices[0][0] == 0 and slices[-1][1] == n
            assert all(slices[i][1] == slices[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in slices]
            assert max(sizes) - min(sizes) <= 1
    assert sharding.stream_assignment(8, 8) == [[i] for i in range(8)]
    assert sharding.stream_assignment(3, 2) == [[0, 1], [2]]
    with pytest.raises(ValueError):
        sharding.balanced_slice(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, result_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m = 96
        taps = oracle.sinc_m2_channelizer(25000.0, m, 9)
        rng = np.random.default_rng(0)                      # same tuner buffer on every rank (H2D broadcast)
        x = sg.interleave(sg.awgn(rng, 48 * 50, 0.1) + sg.tone(2.4e6, 7 * 25000.0 + 800.0, 48 * 50, 0.3))
        res = oracle.Channelizer(taps, m).receive(x)
        lo, hi = sharding.bin_slice(m, rank, world)
        mine = np.stack([oracle.OneChannelOutputProcessor(50000.0, k, float(m)).process(res) for k in range(lo, hi)])
        gathered = [None] * world
        dist.all_gather_object(gathered, (lo, hi, mine))
        ms = sharding.max_over_ranks([10.0 + rank, 5.0 - rank])
        if rank == 0:
            full = np.concatenate([g[2] for g in sorted(gathered, key=lambda g: g[0])])
            want = np.stack([oracle.OneChannelOutputProcessor(50000.0, k, float(m)).process(res) for k in range(m)])
            ok = (full.shape == want.shape and np.array_equal(full, want) and ms == [10.0 + world - 1, 5.0]
                  and [g[:2] for g in sorted(gathered, key=lambda g: g[0])] == [sharding.bin_slice(m, r, world) for r in range(world)])
            open(os.path.join(result_dir, "ok"), "w").write("1" if ok else "0")
    finally:
        dist.destroy_process_group()


def test_bin_sharding_world_size_2_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert open(tmp_path / "ok").read() == "1"
