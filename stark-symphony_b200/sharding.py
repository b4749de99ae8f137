"""Multi-GPU sharding of a proof batch: proofs are independent, so a batch is cut into
contiguous ranges of proof index, one range per rank / GPU, and the only exchange is the gather of the accept
bitmaps (n/8 bytes in total).  There is no collective on the verification path itself."""
from __future__ import annotations

from typing import List, Tuple


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of rank's proofs.  Every shard but the last is a multiple of 32 proofs so that no bitmap word
    is shared between ranks; the remainder goes to the last non-empty shard."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank / world")
    per = ((n + world - 1) // world + 31) // 32 * 32
    begin = min(n, rank * per)
    end = min(n, begin + per)
    return begin, end


def shard_compact(blob, offsets, rank: int, world: int):
    """This rank's slice of a batch held in the compact transport form (witness.compact_stwo): (blob words, offsets rebased to 0) of the
    proofs shard_range gives it.  Views, no copies: records are contiguous and self-describing, so the blob shards by proof index too."""
    n = len(offsets) - 1
    begin, end = shard_range(n, rank, world)
    lo, hi = int(offsets[begin]), int(offsets[end])
    return blob[lo:hi], offsets[begin:end + 1] - offsets[begin]


def shard_words(n: int, world: int) -> int:
    """Bitmap words every rank contributes to the gather (fixed size: all_gather needs equal shapes)."""
    per = ((n + world - 1) // world + 31) // 32 * 32
    return per // 32


def gather_accept_bitmaps(local_bits, n: int, world: int, group=None):
    """All-gather the per-rank accept bitmaps into the bitmap of the whole batch ((n+31)//32 words).
    `local_bits` is a torch int32 tensor of this rank's words (device tensor under NCCL, CPU tensor under gloo)."""
    import torch
    import torch.distributed as dist

    words = shard_words(n, world)
    buf = torch.zeros(words, dtype=torch.int32, device=local_bits.device)
    buf[: local_bits.numel()] = local_bits
    if world == 1 or not dist.is_initialized():
        return buf[: (n + 31) // 32]
    out = torch.empty(words * world, dtype=torch.int32, device=local_bits.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    return out[: (n + 31) // 32]


def expected_accept_count(bitmap_words, n: int) -> int:
    import numpy as np

    bits = np.unpackbits(np.asarray(bitmap_words).astype(np.int32).view(np.uint8), bitorder="little")[:n]
    return int(bits.sum())
