"""Multi-GPU execution of the string operations: keys replicated, haystack windows / radix blocks
partitioned across ranks, and ONE exchange at the narrow end of the tree -- an all-reduce(SUM) of each rank's boolean
block.  u64 wrap-around addition of LWE words IS homomorphic addition, so the reduced block holds the count of ranks
whose share matched (<= world <= 15 = max degree) and one final LUT (x != 0 / x == world) finishes the tree on every
rank.  The collective moves 2049 words (16 KiB); everything else is rank-local.

The sharding logic is backend-agnostic: `execute(program, inputs) -> outputs` is Program.run on an Engine (GPU, NCCL)
or the oracle executor (CPU, gloo) in the tests."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .host import Program


_CACHE: dict = {}


def cached_program(op: str, args: tuple, params: dict) -> Program:
    """programs are recorded once per (op, shape, parameter set) and stay bound to the context that first ran them"""
    key = (op, tuple(args), tuple(sorted(params.items())))
    if key not in _CACHE:
        _CACHE[key] = Program(op, args, params=params)
    return _CACHE[key]


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """contiguous, balanced partition of n_items (first ranks get the remainder)"""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_reduce_lwe(block: np.ndarray, device: str | None = None) -> np.ndarray:
    """sum of LWE ciphertexts over ranks == homomorphic addition (wrapping u64 via two's-complement int64)"""
    t = torch.from_numpy(block.view(np.int64).copy())
    if device:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy().view(np.uint64)


def sharded_contains(execute, params: dict, hay: np.ndarray, pat: np.ndarray, hay_len: int, pat_len: int,
                     rank: int, world: int, device: str | None = None) -> np.ndarray:
    """contains(hay, pat) with the windows split over `world` ranks.  Every rank returns the same boolean block."""
    n_win = hay_len - pat_len + 1
    if n_win <= 0 or pat_len == 0 or world == 1:   # nothing to split: the plain single-GPU program
        return cached_program("string_contains", (hay_len, pat_len), params).pipe(execute, np.concatenate([hay, pat]))[0]
    active = min(world, n_win)                      # more ranks than windows: the surplus ranks contribute a zero block
    inputs = np.concatenate([hay, pat])
    if rank < active:
        w0, w1 = shard_range(n_win, rank, active)
        mine = execute(cached_program("string_contains_windows", (hay_len, pat_len, w0, w1), params), inputs)[0]
    else:
        mine = np.zeros(inputs.shape[1], dtype=np.uint64)
    total = all_reduce_lwe(mine, device) if world > 1 else mine
    fin = cached_program("bool_sum_finish", (active, 0), params)
    return execute(fin, total[None, :])[0]


def sharded_eq(execute, params: dict, a: np.ndarray, b: np.ndarray, n_chars: int, rank: int, world: int,
               device: str | None = None) -> np.ndarray:
    """eq of two equal-length strings with the chars split over ranks: per-rank eq of its slice, all-reduce, x == active."""
    if world == 1 and n_chars:
        return execute(cached_program("string_eq", (n_chars, n_chars), params), np.concatenate([a, b]))[0]
    active = max(1, min(world, n_chars))
    if rank < active and n_chars:
        c0, c1 = shard_range(n_chars, rank, active)
        ins = np.concatenate([a[4 * c0:4 * c1], b[4 * c0:4 * c1]])
        mine = execute(cached_program("string_eq", (c1 - c0, c1 - c0), params), ins)[0]
    elif n_chars == 0:
        return execute(cached_program("string_eq", (0, 0), params), np.zeros((0, a.shape[1] if a.ndim == 2 else 1), dtype=np.uint64))[0]
    else:
        mine = np.zeros(a.shape[1], dtype=np.uint64)
    total = all_reduce_lwe(mine, device) if world > 1 else mine
    return execute(cached_program("bool_sum_finish", (active, 1), params), total[None, :])[0]
