"""Frame-pair sharding for multi-GPU interpolation.

Every (frame pair, t) unit of ``convert.py:120-130`` is independent, so the path shards with no
data-path collective: rank r of R owns a contiguous range of pairs (pair i needs frames i and
i+1, so consecutive shards share one boundary frame) and all timesteps of its pairs (the
t-independent Flow U-Net, model.py:33-35, is then computed once per pair on one GPU).
Pure host logic: no CUDA, no torch.distributed calls -- callers pass rank / world size.
"""
from __future__ import annotations

from typing import List, Tuple


def pair_range(n_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) range of pair indices owned by ``rank``; pair i interpolates frames (i, i+1).

    Shard sizes differ by at most one; earlier ranks take the larger shards.
    """
    if n_frames < 0 or world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad shard request: n_frames={n_frames} rank={rank} world={world}")
    n_pairs = max(n_frames - 1, 0)
    base, extra = divmod(n_pairs, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def frame_range(n_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) range of source frames ``rank`` must hold (its pairs plus the shared boundary frame)."""
    lo, hi = pair_range(n_frames, rank, world)
    return (lo, hi + 1) if hi > lo else (lo, lo)


def output_index(pair: int, k: int, sf: int) -> int:
    """Position of interpolated frame k (1..sf) of ``pair`` in the output sequence, which
    interleaves originals and interpolated frames like convert.py:124-142:
    original i sits at i*(sf+1), its interpolations at i*(sf+1)+k."""
    if not 1 <= k <= sf:
        raise ValueError(f"k must be in 1..{sf}, got {k}")
    return pair * (sf + 1) + k


def timesteps(sf: int) -> List[float]:
    """t = i/(sf+1), i = 1..sf (convert.py:127-129)."""
    return [i / (sf + 1) for i in range(1, sf + 1)]


def plan(n_frames: int, world: int, sf: int = 1):
    """[(rank, pair, k, t, output_index)] for the whole job, rank-major -- the union over ranks is
    exactly every (pair, k) once; used by the tests and by bench.py."""
    ts = timesteps(sf)
    out = []
    for r in range(world):
        lo, hi = pair_range(n_frames, r, world)
        for p in range(lo, hi):
            for k, t in enumerate(ts, 1):
                out.append((r, p, k, t, output_index(p, k, sf)))
    return out
