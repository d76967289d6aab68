"""Multi-GPU plumbing of the decode path: one process per GPU, shots sharded across ranks, DEM tables replicated,
and ONE collective — an all-reduce (sum) of the per-rank counters {shots, logical errors, low confidence}
(the reference shards the same way across CPU threads, utils.h:88-92, and by --shot-range-*).
torch.distributed is the transport (NCCL over NVLink on GPUs, gloo in the CPU tests); no search data crosses ranks.
"""
from __future__ import annotations

import numpy as np


def shard_chunks(num_shots: int, rank: int, world: int, chunk: int = 4096):
    """Interleaved fixed-size chunks of [0, num_shots): rank r owns chunks r, r+world, ...  Per-shot cost is
    heavy-tailed, so interleaving balances better than `world` contiguous ranges.  Returns [(begin, end), ...]."""
    if not (0 <= rank < world) or chunk <= 0:
        raise ValueError("need 0 <= rank < world and chunk > 0")
    out = []
    for k, begin in enumerate(range(0, num_shots, chunk)):
        if k % world == rank:
            out.append((begin, min(begin + chunk, num_shots)))
    return out


def shard_rows(array: np.ndarray, rank: int, world: int, chunk: int = 4096) -> np.ndarray:
    """This rank's rows of a [shots, ...] array, in shot order."""
    parts = [array[b:e] for b, e in shard_chunks(len(array), rank, world, chunk)]
    return np.concatenate(parts, axis=0) if parts else array[:0]


def count_errors(pred_obs: np.ndarray, true_obs: np.ndarray, low_confidence: np.ndarray):
    """(shots, logical errors, low confidence) of one shard with the CLI's accounting (tesseract_main.cc:597-601):
    a low-confidence shot is never a logical error."""
    low = np.asarray(low_confidence).astype(bool)
    wrong = (np.asarray(pred_obs) != np.asarray(true_obs)).reshape(len(low), -1).any(axis=1) & ~low
    return int(len(low)), int(wrong.sum()), int(low.sum())


def all_reduce_counts(counts, device=None):
    """Sum of the three counters over all ranks (identity when torch.distributed is not initialised)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(counts), dtype=torch.int64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t)
    return tuple(int(v) for v in t.tolist())


def decode_sharded(decode_b8, dets_b8: np.ndarray, true_obs: np.ndarray, rank: int, world: int, chunk: int = 4096, device=None):
    """Decodes this rank's shard with `decode_b8(dets) -> {"obs", "low_confidence", ...}` and returns the
    whole-job (shots, logical errors, low confidence) after the all-reduce, plus this rank's predictions."""
    mine = shard_rows(dets_b8, rank, world, chunk)
    out = decode_b8(mine)
    local = count_errors(out["obs"], shard_rows(true_obs, rank, world, chunk), out["low_confidence"])
    return all_reduce_counts(local, device=device), out
