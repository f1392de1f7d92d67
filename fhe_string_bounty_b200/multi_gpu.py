"""Multi-GPU execution of the string operations: keys replicated, haystack windows / radix blocks
partitioned across ranks, and ONE exchange at the narrow end of the tree -- for the boolean trees (eq, contains) an
all-reduce(SUM) of each rank's boolean block; for the lexicographic comparisons an all-gather of each rank's sign block
followed by the pairwise sign tree; for find an all-gather of each rank's (found, first index) followed by a two-level
first-rank selection; for the elementwise case conversions an all-gather of the converted chars and nothing else.  u64 wrap-around addition of LWE words IS homomorphic addition, so the reduced block holds the count of ranks
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


def all_gather_lwe(blocks: np.ndarray, device: str | None = None) -> np.ndarray:
    """every rank's (m, lwe_len) block array -> (world, m, lwe_len) on every rank (all ranks pass the same m)"""
    t = torch.from_numpy(np.ascontiguousarray(blocks).view(np.int64).copy())
    if device:
        t = t.to(device)
    world = dist.get_world_size()
    out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)   # concatenated along dim 0
    dist.all_gather_into_tensor(out, t)
    return out.cpu().numpy().view(np.uint64).reshape((world,) + tuple(t.shape))


_CMP = {"lt": (1, 0), "le": (1, 1), "gt": (0, 0), "ge": (0, 1)}     # (want_less, or_equal) of strings.h cmp()


def sharded_compare(execute, params: dict, op: str, a: np.ndarray, b: np.ndarray, n_chars: int, rank: int, world: int,
                    device: str | None = None) -> np.ndarray:
    """lt / le / gt / ge of two equal-length strings with the chars split over ranks (lexicographic tree): every rank
    reduces its char range to one sign block (comparator.rs:389-464), the sign blocks are all-gathered (world x 16 KiB) and every
    rank finishes with the pairwise sign tree + map_sign_result (ceil(log2(world)) + 1 levels of at most world/2 PBS)."""
    want_less, or_equal = _CMP[op]
    active = max(1, min(world, n_chars))
    if world == 1 or n_chars == 0:
        return execute(cached_program("string_" + op, (n_chars, n_chars), params), np.concatenate([a, b]))[0]
    if rank < active:
        c0, c1 = shard_range(n_chars, rank, active)
        ins = np.concatenate([a[4 * c0:4 * c1], b[4 * c0:4 * c1]])
        mine = execute(cached_program("string_cmp_sign", (c1 - c0, c1 - c0), params), ins)
    else:
        mine = np.zeros((1, a.shape[1]), dtype=np.uint64)
    signs = all_gather_lwe(mine, device)[:active, 0]
    # char 0 is the most significant: the sign tree wants the least significant range first
    return execute(cached_program("signs_finish", (active, want_less, or_equal), params), np.ascontiguousarray(signs[::-1]))[0]


def sharded_case(execute, params: dict, op: str, s: np.ndarray, n_chars: int, rank: int, world: int,
                 device: str | None = None, gather: bool = True) -> np.ndarray:
    """to_lowercase / to_uppercase with the chars split over ranks: elementwise, so there is no exchange on the path.  With
    gather=True every rank returns the whole converted string (one all-gather of the converted blocks, 16 KiB per block -- for a
    1024-char string that is 67 MB and costs more than the conversion); with gather=False a rank returns its own chars
    [shard_range(n_chars, rank, world)] only ("no exchange; optional all-gather")."""
    if world == 1 or n_chars == 0:
        return execute(cached_program("string_" + op, (n_chars,), params), s)
    active = min(world, n_chars)
    per = -(-n_chars // active)                                  # padded share, in chars
    c0 = c1 = 0
    if rank < active:
        c0, c1 = shard_range(n_chars, rank, active)
    conv = execute(cached_program("string_" + op, (c1 - c0,), params), s[4 * c0:4 * c1]) if c1 > c0 else np.zeros((0, s.shape[1]), dtype=np.uint64)
    if not gather:
        return conv
    mine = np.zeros((4 * per, s.shape[1]), dtype=np.uint64)
    mine[:4 * (c1 - c0)] = conv
    parts = all_gather_lwe(mine, device)
    out = []
    for r in range(active):
        r0, r1 = shard_range(n_chars, r, active)
        out.append(parts[r, :4 * (r1 - r0)])
    return np.concatenate(out)


def sharded_find(execute, params: dict, hay: np.ndarray, pat: np.ndarray, hay_len: int, pat_len: int, rank: int, world: int,
                 device: str | None = None) -> np.ndarray:
    """find(hay, pat) with the windows split over ranks: every rank finds the first match inside its window range (reported as a
    global index), the (found, index) blocks are all-gathered and every rank selects the lowest rank that found (two PBS levels).
    Returns [found, index blocks...] like the single-GPU program."""
    n_win = hay_len - pat_len + 1
    inputs = np.concatenate([hay, pat])
    if n_win <= 1 or pat_len == 0 or world == 1:
        return execute(cached_program("string_find", (hay_len, pat_len), params), inputs)
    total_mod = params["msg_mod"] * params["carry_mod"]
    active = min(world, n_win, total_mod // 2)
    if rank < active:
        w0, w1 = shard_range(n_win, rank, active)
        mine = execute(cached_program("string_find_windows", (hay_len, pat_len, w0, w1), params), inputs)
    else:
        nb = 1
        while (1 << (2 * nb)) < n_win:
            nb += 1
        mine = np.zeros((1 + nb, inputs.shape[1]), dtype=np.uint64)
    parts = all_gather_lwe(mine, device)[:active]
    nb = parts.shape[1] - 1
    return execute(cached_program("find_combine", (active, nb), params), np.ascontiguousarray(parts.reshape(active * (1 + nb), -1)))
