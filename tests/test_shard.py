"""Host-side multi-GPU logic on CPU: the byte-balanced partition and the one optional exchange (an all-gather of
per-image checksums) over gloo with world_size 2. The per-rank decoder here is the oracle (tests may use it);
on the GPU box the same code path runs with the CUDA decoders (bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest

from gamut_b200 import shard


def test_partition_properties():
    rng = np.random.default_rng(0)
    for world in (1, 2, 3, 8):
        for n in (0, 1, 5, 64, 257):
            sizes = rng.integers(1, 5_000_000, n).tolist()
            parts = shard.partition(sizes, world)
            assert len(parts) == world
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            assert all(s <= e for s, e in parts)
            if n >= 8 * world:
                tot = sum(sizes)
                worst = max(sum(sizes[s:e]) for s, e in parts)
                assert worst <= tot / world + max(sizes)       # balanced to within one image
    # equal sizes => equal counts
    assert shard.partition([7] * 1024, 8) == [(i * 128, (i + 1) * 128) for i in range(8)]


def test_checksum_sensitivity():
    a = np.arange(1000, dtype=np.uint8)
    b = a.copy(); b[[3, 4]] = b[[4, 3]]
    assert shard.checksum64(a) != shard.checksum64(b)
    assert shard.checksum64(a) == shard.checksum64(a.copy())
    assert shard.checksum64(a[:999]) != shard.checksum64(a)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, files, q):
    import torch.distributed as dist
    from oracle import pyoracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ranges = shard.partition([len(f) for f in files], world)
        s, e = ranges[rank]
        local = []
        for f in files[s:e]:
            px, _ = pyoracle.png_load(f, 0, 0)
            local.append(shard.checksum64(px))
        allsums = shard.gather_checksums(local, ranges)
        q.put((rank, (s, e), allsums.tolist()))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_batch():
    import torch.multiprocessing as mp
    from oracle import pyoracle
    from pngwriter import write_png
    from test_oracle_png import synth
    files = [write_png(synth(16 + 3 * i, 9 + i, 3 + (i % 2), 8, i), 2 + 4 * (i % 2), 8, filters=i % 5) for i in range(11)]
    expect = []
    for f in files:
        px, _ = pyoracle.png_load(f, 0, 0)
        expect.append(shard.checksum64(px))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, files, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in procs]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    res.sort()
    (r0, rg0, s0), (r1, rg1, s1) = res
    assert rg0[0] == 0 and rg0[1] == rg1[0] and rg1[1] == len(files) and rg0[1] > 0 and rg1[1] > rg1[0]
    assert s0 == expect and s1 == expect
