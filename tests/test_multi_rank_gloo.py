"""world_size-2 (and 3) gloo test of the multi-GPU sharding path on CPU: window / char partition, all-reduce(SUM) of the
per-rank boolean LWE blocks (wrapping u64 add == homomorphic add) + final LUT for eq / contains; all-gather of sign blocks + sign
tree for lt / le / gt / ge; all-gather of converted chars for the case conversions; all-gather of (found, index) + first-rank
selection for find.  Programs are executed with the CPU oracle;
the sharding + collective code is the product's (fhe_string_bounty_b200/multi_gpu.py)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from oracle import radix as R
    from fhe_string_bounty_b200 import multi_gpu as MG
    from helpers import engine_params
    p = O.params("toy")
    ck = O.ClientKey(p, 0xB200)          # same seeds on every rank: keys are replicated
    sk = O.ServerKey(ck, 0xB201)
    params = engine_params(p)
    execute = lambda prog, ins: R.run_program(prog.ir(), sk, ins)
    results = []
    for hay, pat in [(b"the quick brown fox", b"brown"), (b"the quick brown fox", b"browm"), (b"aaab", b"ab"), (b"ab", b"ab")]:
        h, q = R.encrypt_string(ck, hay), R.encrypt_string(ck, pat)
        out = MG.sharded_contains(execute, params, h, q, len(hay), len(pat), rank, world)
        results.append(("contains", hay, pat, ck.decrypt_message_and_carry(out), int(pat in hay)))
    for a, b in [(b"abcdefg", b"abcdefg"), (b"abcdefg", b"abcdefh"), (b"x", b"x")]:
        out = MG.sharded_eq(execute, params, R.encrypt_string(ck, a), R.encrypt_string(ck, b), len(a), rank, world)
        results.append(("eq", a, b, ck.decrypt_message_and_carry(out), int(a == b)))
    for a, b in [(b"abcdefg", b"abcdefg"), (b"abcdefg", b"abcdefh"), (b"bbcdefg", b"abcdefz"), (b"aaaaaab", b"aaaaaaa"), (b"q", b"r")]:
        for op, w in (("lt", a < b), ("le", a <= b), ("gt", a > b), ("ge", a >= b)):
            out = MG.sharded_compare(execute, params, op, R.encrypt_string(ck, a), R.encrypt_string(ck, b), len(a), rank, world)
            results.append((op, a, b, ck.decrypt_message_and_carry(out), int(w)))
    for s_ in [b"Hello Zama, how is it going?", b"aZ"]:
        out = MG.sharded_case(execute, params, "to_uppercase", R.encrypt_string(ck, s_), len(s_), rank, world)
        results.append(("upper", s_, b"", R.decrypt_string(ck, out), s_.upper()))
        out = MG.sharded_case(execute, params, "to_lowercase", R.encrypt_string(ck, s_), len(s_), rank, world)
        results.append(("lower", s_, b"", R.decrypt_string(ck, out), s_.lower()))
        c0, c1 = MG.shard_range(len(s_), rank, min(world, len(s_))) if rank < min(world, len(s_)) else (0, 0)
        out = MG.sharded_case(execute, params, "to_lowercase", R.encrypt_string(ck, s_), len(s_), rank, world, gather=False)
        results.append(("lower-own-share", s_, b"", R.decrypt_string(ck, out), s_.lower()[c0:c1]))
    for hay, pat in [(b"the quick brown fox", b"quick"), (b"abcabcabc", b"abc"), (b"abcabcabc", b"cab"), (b"abcabcabd", b"abd"), (b"abcabc", b"xyz")]:
        out = MG.sharded_find(execute, params, R.encrypt_string(ck, hay), R.encrypt_string(ck, pat), len(hay), len(pat), rank, world)
        pos = hay.find(pat)
        results.append(("find", hay, pat, (ck.decrypt_message_and_carry(out[0]), R.decrypt_radix(ck, out[1:])), (int(pos >= 0), max(pos, 0))))
    assert MG.shard_range(241, 0, 8) == (0, 31) and MG.shard_range(241, 7, 8) == (211, 241)
    ret[rank] = results
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharded_string_ops_gloo(world):
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == world
    for rank in range(world):
        for op, a, b, got, want in ret[rank]:
            assert got == want, (rank, op, a, b, got, want)
