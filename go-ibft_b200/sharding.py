"""Multi-GPU sharding of the signature batch: contiguous 32-aligned index ranges per rank, validator tables
replicated, ONE all-gather of the pass/fail bitmap words (NCCL over NVLink on GPUs; gloo in the CPU tests), after which
every rank reduces quorum on the complete bitmap.  Works with any torch.distributed backend."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int, rank: int) -> tuple[int, int]:
    """[lo, hi) of `rank`: contiguous, disjoint, covering [0, n), every boundary except n a multiple of 32 (one bitmap word
    never straddles two ranks).  Trailing ranks may get an EMPTY shard (lo == hi == n): skip the verify call for those."""
    words = (n + 31) // 32
    per = (words + world - 1) // world
    lo = min(rank * per * 32, n)
    hi = min((rank + 1) * per * 32, n)
    return lo, hi


def shard_words(n: int, world: int) -> int:
    """bitmap words every rank contributes to the all-gather (equal on all ranks; trailing ranks pad with zeros)."""
    return ((n + 31) // 32 + world - 1) // world


def all_gather_bitmap(local_words: torch.Tensor, n: int, world: int) -> torch.Tensor:
    """local_words: int32 tensor of shard_words(n, world) words (this rank's slice, zero padded).  Returns the complete
    bitmap of (n+31)//32 words on every rank.  One collective."""
    per = shard_words(n, world)
    assert local_words.numel() == per
    if world == 1:
        return local_words[: (n + 31) // 32]
    full = torch.empty(per * world, dtype=local_words.dtype, device=local_words.device)
    dist.all_gather_into_tensor(full, local_words.contiguous())
    return full[: (n + 31) // 32]
