"""Multi-GPU sharding of a texture batch: independent images, no exchange.

One process per GPU (torchrun); rank r transcodes the images the plan assigns to it with the same
C-ABI calls a single-GPU user makes.  There is no collective on the data path -- results stay with
the owning rank (or are written out by it); torch.distributed is only used by callers that want a
barrier or a max-over-ranks time.  The plan is a pure function of the batch description, so every
rank computes it locally and they all agree without communicating.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Sequence, Tuple


def image_cost(nblocks: int, is_etc1s: bool) -> float:
    """Relative device time of one image.  ETC1S slices are dominated by the serial entropy decode
    (one warp per slice), UASTC blocks are independent; the constant only has to rank images."""
    return float(nblocks) * (8.0 if is_etc1s else 1.0)


def plan_shards(costs: Sequence[float], world: int) -> List[List[int]]:
    """Assigns image i to a rank.  Equal costs -> i mod world (keeps mip chains of consecutive images on
    neighbouring ranks); otherwise greedy longest-processing-time-first, ties broken by index so that every
    rank derives the identical plan."""
    if world < 1:
        raise ValueError("world must be >= 1")
    n = len(costs)
    shards: List[List[int]] = [[] for _ in range(world)]
    if n == 0:
        return shards
    if all(c == costs[0] for c in costs):
        for i in range(n):
            shards[i % world].append(i)
        return shards
    load = [0.0] * world
    for i in sorted(range(n), key=lambda i: (-costs[i], i)):
        r = min(range(world), key=lambda r: (load[r], r))
        shards[r].append(i)
        load[r] += costs[i]
    for s in shards:
        s.sort()
    return shards


def transcode_batch(files: Sequence[bytes], target: int, rank: int, world: int,
                    read_to: Callable[[int, bytes], Tuple[object, list]] | None = None,
                    costs: Sequence[float] | None = None) -> Dict[int, list]:
    """Transcodes this rank's share of `files` (.basis payloads) to `target`; returns {image index: [Image, ...]}.
    `read_to(target, buf)` defaults to the CUDA path (basisu_rs_b200._read_to); tests inject a checker."""
    if read_to is None:
        import basisu_rs_b200 as b
        read_to = b._read_to
    if costs is None:
        costs = [float(len(f)) for f in files]
    mine = plan_shards(costs, world)[rank]
    return {i: read_to(target, files[i])[1] for i in mine}
