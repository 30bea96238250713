"""Multi-process (world_size 2, gloo, CPU) test of the sharding and gather logic of the N>1 path.
The per-rank compute here is the CPU oracle (test infrastructure): what is under test is the partitioning (stream ranges,
frame ranges with halo, voice ranges) and the final gather / bus reduction, which are the same code on NCCL."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    pkg = g.load_package()
    orc = g.load_oracle()
    import cpp_audio_b200.workloads as wl
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        W, hop = 4096, 256
        # (1) independent streams: 5 streams over 2 ranks
        x = wl.c4_streams(5, 0.3)
        s0, s1 = pkg.shard_range(5, rank, world)
        r = orc.stft_analyze(x[s0:s1], W, hop, max_peaks=32)
        local = torch.from_numpy(r["peak_freq"].reshape(-1, 32))
        allf = pkg.gather_rows(local)
        # (2) one long stream split by frame ranges with halo
        y = wl.c2_signal(1.0)
        f0, f1, a, b = pkg.frame_shard(len(y), W, hop, rank, world)
        rr = orc.stft_analyze(y[a:b], W, hop, max_peaks=32)
        assert rr["frames"] == f1 - f0
        allg = pkg.gather_rows(torch.from_numpy(rr["peak_freq"].reshape(-1, 32)))
        # (3) voices: mixed bus is the sum of the per-rank buses
        bank = wl.c3_bank(6, 8)
        v0, v1 = pkg.shard_range(6, rank, world)
        bus = orc.oscbank_render(bank["f0_hz"][v0:v1], bank["volume"][v0:v1], bank["partial_volume"][v0:v1],
                                 bank["partial_phase"][v0:v1], 2000, mix=2)
        bus = pkg.reduce_bus(torch.from_numpy(bus))
        if rank == 0:
            q.put((allf.numpy(), allg.numpy(), bus.numpy()))
    finally:
        dist.destroy_process_group()


def test_shard_range_covers_everything(pkg):
    for n in (0, 1, 7, 8, 9, 1024):
        for world in (1, 2, 3, 8):
            parts = [pkg.shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_frame_shard_halo(pkg):
    n, W, hop = 2880000, 4096, 256
    spans = [pkg.frame_shard(n, W, hop, r, 8) for r in range(8)]
    assert sum(f1 - f0 for f0, f1, _, _ in spans) == 11235
    for f0, f1, a, b in spans:
        assert a == f0 * hop and b == (f1 - 1) * hop + W and b <= n
    # neighbouring shards overlap by exactly window - hop samples
    assert spans[0][3] - spans[1][2] == W - hop


def test_two_ranks_match_single_process(pkg, orc, wl):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    allf, allg, bus = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x = wl.c4_streams(5, 0.3)
    ref = orc.stft_analyze(x, 4096, 256, max_peaks=32)
    assert np.array_equal(allf, ref["peak_freq"].reshape(-1, 32))
    y = wl.c2_signal(1.0)
    ref2 = orc.stft_analyze(y, 4096, 256, max_peaks=32)
    assert np.array_equal(allg, ref2["peak_freq"].reshape(-1, 32))
    bank = wl.c3_bank(6, 8)
    ref3 = orc.oscbank_render(bank["f0_hz"], bank["volume"], bank["partial_volume"], bank["partial_phase"], 2000, mix=2)
    assert np.abs(bus - ref3).max() < 1e-12
