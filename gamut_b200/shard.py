"""Multi-GPU sharding of a batch of independent images.

Images share no state (every reference `loadProc` call touches only its own `Image`, plugin.d:30), so a batch is
cut into contiguous ranges of image indices, one range per rank, balanced by compressed bytes (PNG/JPEG sizes vary);
every rank decodes its own range with no collective on the data path. The only exchanges are optional: one
all-gather of per-image 64-bit checksums (or of the decoded sizes) for verification/reporting. One process per GPU
(`torchrun`), `torch.distributed` with NCCL on GPUs; the same code runs over gloo on CPU (tests/test_shard.py)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def partition(sizes: Sequence[int], world: int) -> List[Tuple[int, int]]:
    """Contiguous [start, end) ranges, one per rank, whose byte totals are as close as possible to total/world
    (greedy on the running prefix: a range is closed at the index whose prefix sum is nearest to its target)."""
    n = len(sizes)
    if world <= 0:
        raise ValueError("world must be positive")
    pref = np.concatenate([[0], np.cumsum(np.asarray(sizes, dtype=np.int64))])
    total = int(pref[-1])
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        k = int(np.searchsorted(pref, target))
        if k > 0 and abs(pref[k - 1] - target) <= abs(pref[min(k, n)] - target):
            k -= 1
        k = min(max(k, cuts[-1]), n)
        cuts.append(k)
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def my_range(sizes: Sequence[int], rank: int, world: int) -> Tuple[int, int]:
    return partition(sizes, world)[rank]


def checksum64(a: np.ndarray) -> int:
    """Order-sensitive 64-bit checksum of a decoded image (FNV-1a over 8-byte words, vectorised per lane)."""
    b = np.ascontiguousarray(a).view(np.uint8).ravel()
    pad = (-b.size) % 8
    if pad:
        b = np.concatenate([b, np.zeros(pad, np.uint8)])
    w = b.view(np.uint64)
    idx = np.arange(1, w.size + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        h = np.bitwise_xor.reduce((w + idx * np.uint64(0x9E3779B97F4A7C15)) * np.uint64(0x100000001B3)) if w.size else np.uint64(0)
    return int(h) ^ (b.size - pad)


def gather_checksums(local: Sequence[int], ranges: Sequence[Tuple[int, int]], device=None) -> np.ndarray:
    """All ranks receive the checksums of the whole batch, in image order. `local` holds this rank's range.
    One all_gather of a fixed-size int64 tensor (ranges are padded to the longest)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    n = ranges[-1][1]
    out = np.zeros(n, np.uint64)
    if world == 1:
        out[ranges[0][0]:ranges[0][1]] = np.asarray(local, dtype=np.uint64)
        return out
    longest = max(e - s for s, e in ranges)
    buf = np.zeros(longest, np.int64)
    buf[:len(local)] = np.asarray(local, dtype=np.uint64).view(np.int64)
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    for r, (s, e) in enumerate(ranges):
        out[s:e] = parts[r].cpu().numpy()[:e - s].view(np.uint64)
    assert ranges[rank][1] - ranges[rank][0] == len(local)
    return out
