"""GPU tests (-m gpu) of the device-resident program path and of the engine's own exchange kernel (csrc/exchange.cu):
  * tfhe_b200_program_run_device == tfhe_b200_program_run, word for word;
  * clear-operand programs (trivial ciphertexts folded on the host like trivial_pbs_assign, shortint/server_key/mod.rs:763-791) and the
    scalar radix comparisons (radix_parallel/scalar_comparison.rs:366-458, comparator.rs:474-502) on real keys;
  * find / rfind on a haystack with more than 420 windows and dense matches (the carry of the block-prefix sums);
  * every sharded string operation with two ranks driven from one process on ONE GPU: the ranks' exchange buffers are attached by
    pointer, each rank runs on its own stream, and the publish / wait / pull code of the exchange kernel runs for both ranks in one
    cooperative launch (ranks that share a GPU must not spin in separate launches) -- decrypted results against clear text
    (scalar_comparison.rs:147-240 for the boolean trees, comparator.rs:257-279 for the sign tree).
The two-GPU NCCL / CUDA-IPC variant is tests/test_gpu_multi_rank.py (skipped on a one-GPU box)."""
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from fhe_string_bounty_b200.host import Program
from helpers import engine_params

pytestmark = pytest.mark.gpu


def _engine(p, sk):
    import fhe_string_bounty_b200 as F
    e = F.Engine(engine_params(p))
    e.upload_ksk(sk.ksk)
    e.upload_bsk_std(sk.bsk)
    return e


@pytest.fixture(scope="module")
def eng(keys_2_2):
    p, ck, sk = keys_2_2
    e = _engine(p, sk)
    yield e
    e.close()


def _bool(ck, ct):
    return ck.decrypt_message_and_carry(ct)


def test_program_run_device_matches_host_path(orc, keys_2_2, eng):
    import torch
    from oracle import radix as R
    from fhe_string_bounty_b200.multi_gpu import DeviceComm
    p, ck, sk = keys_2_2
    comm = DeviceComm(eng, rank=0, world=1)
    a, b = b"device path", b"device pAth"
    ins = np.concatenate([R.encrypt_string(ck, a), R.encrypt_string(ck, b)])
    for op in ("string_eq", "string_lt", "string_find", "string_eq_ignore_case"):
        P = Program(op, (len(a), len(b)), params=engine_params(p))
        host = P.run(eng, ins)
        with torch.cuda.stream(torch.cuda.Stream()):
            dev = comm.to_host(comm.run(P, ins))
        assert np.array_equal(host, dev), op
        assert P.last_ms() > 0
    comm.close()


def test_clear_operand_programs_gpu(orc, keys_2_2, eng):
    """second operand = clear string: its blocks are trivial ciphertexts, every LUT whose operand is entirely clear is evaluated on the
    host (n_trivial_pbs), the mixed ones run with the clear part folded into the body"""
    from oracle import radix as R
    p, ck, sk = keys_2_2
    hay = b"the quick brown fox jumps"
    enc = R.encrypt_string(ck, hay)
    for pat in (b"brown", b"browm", b"the", b"jumps", b"", b"x" * 30):
        for op, want in (("contains", pat in hay), ("starts_with", hay.startswith(pat)), ("ends_with", hay.endswith(pat)),
                         ("eq", hay == pat), ("lt", hay < pat), ("ge", hay >= pat)):
            P = Program("string_" + op, (len(hay),), clear=pat, params=engine_params(p))
            out = P.run(eng, enc)
            assert _bool(ck, out[0]) == int(want), (op, pat)
        P = Program("string_find", (len(hay),), clear=pat, params=engine_params(p))
        out = P.run(eng, enc)
        f = hay.find(pat)
        assert _bool(ck, out[0]) == int(f >= 0) and R.decrypt_radix(ck, out[1:]) == max(f, 0), pat
    assert _bool(ck, Program("string_eq", (len(hay),), clear=hay, params=engine_params(p)).run(eng, enc)[0]) == 1


def test_radix_scalar_comparisons_gpu(orc, keys_2_2, eng):
    from oracle import radix as R
    p, ck, sk = keys_2_2
    rng = np.random.default_rng(77)
    nb = 16
    for trial in range(4):
        x = int(rng.integers(0, 2**32))
        for y in (x, int(rng.integers(0, 2**32)), 0, 2**32 - 1, x ^ 1, 2**34 + 5):
            ins = np.stack(R.encrypt_radix(ck, x, nb))
            for op, w in (("scalar_eq", x == y), ("scalar_lt", x < y), ("scalar_gt", x > y)):
                out = Program("radix_" + op, (nb, y), params=engine_params(p)).run(eng, ins)
                assert _bool(ck, out[0]) == int(w), (op, x, y)


def test_default_comparisons_dirty_carries_gpu(orc, keys_2_2, eng):
    """tests_cases_comparisons.rs:81-97: raise the degree of both operands with unchecked_add (carries not empty), then compare with the
    default (propagating) forms; also full_propagate on blocks filled up to the whole message space"""
    from oracle import radix as R
    p, ck, sk = keys_2_2
    rng = np.random.default_rng(91)
    nb = 16
    for trial in range(3):
        xs = [int(rng.integers(0, 2**32)) for _ in range(2)]
        ys = xs if trial == 0 else [int(rng.integers(0, 2**32)) for _ in range(2)]
        cx = np.stack(R.encrypt_radix(ck, xs[0], nb)) + np.stack(R.encrypt_radix(ck, xs[1], nb))     # wrapping u64 add == unchecked_add
        cy = np.stack(R.encrypt_radix(ck, ys[0], nb)) + np.stack(R.encrypt_radix(ck, ys[1], nb))
        vx, vy = sum(xs) % 2**32, sum(ys) % 2**32
        out = Program("radix_full_propagate", (nb, 6), params=engine_params(p)).run(eng, cx)
        assert R.decrypt_radix(ck, out) == vx
        for op, w in (("eq", vx == vy), ("ne", vx != vy), ("lt", vx < vy), ("le", vx <= vy), ("gt", vx > vy), ("ge", vx >= vy)):
            out = Program("radix_default_" + op, (nb, 6), params=engine_params(p)).run(eng, np.concatenate([cx, cy]))
            assert _bool(ck, out[0]) == int(w), (op, vx, vy)
    five = [np.stack(R.encrypt_radix(ck, int(v), nb)) for v in rng.integers(0, 2**32, size=5)]     # degree 15
    vals = [R.decrypt_radix(ck, f) for f in five]
    out = Program("radix_full_propagate", (nb, 15), params=engine_params(p)).run(eng, sum(five[1:], five[0]))
    assert R.decrypt_radix(ck, out) == sum(vals) % 2**32


def test_find_more_than_420_windows_gpu(orc, keys_2_2, eng):
    from oracle import radix as R
    p, ck, sk = keys_2_2
    hay, pat = b"ab" * 230, b"ab"          # 459 windows = 33 blocks of 14, a match in every block
    ins = np.concatenate([R.encrypt_string(ck, hay), R.encrypt_string(ck, pat)])
    for op, want in (("string_find", hay.find(pat)), ("string_rfind", hay.rfind(pat))):
        out = Program(op, (len(hay), len(pat)), params=engine_params(p)).run(eng, ins)
        assert _bool(ck, out[0]) == 1 and R.decrypt_radix(ck, out[1:]) == want, op


def test_sharded_ops_two_ranks_one_gpu(orc, keys_2_2):
    """world = 2 on one GPU: two engines, two threads, two streams; exchange through the engine's peer-memory kernel (group launch)"""
    import torch
    from oracle import radix as R
    from fhe_string_bounty_b200 import multi_gpu as MG
    p, ck, sk = keys_2_2
    params = engine_params(p)
    engines = [_engine(p, sk) for _ in range(2)]
    comms = MG.DeviceComm.local_group(engines)
    for c in comms:
        c.min_shard_width = 0                        # shard even these small trees: the test is about the sharded path
    streams = [torch.cuda.Stream() for _ in comms]

    def both(fn):
        def one(r):
            with torch.cuda.stream(streams[r]):
                return fn(comms[r])
        with ThreadPoolExecutor(2) as pool:
            res = list(pool.map(one, range(2)))
        return res

    for hay, pat in ((b"the quick brown fox", b"brown"), (b"the quick brown fox", b"browm"), (b"aaab", b"ab")):
        h, q = R.encrypt_string(ck, hay), R.encrypt_string(ck, pat)
        for out in both(lambda c: MG.sharded_contains(c, params, h, q, len(hay), len(pat))):
            assert _bool(ck, out) == int(pat in hay), (hay, pat)
        pos = hay.find(pat)
        for out in both(lambda c: MG.sharded_find(c, params, h, q, len(hay), len(pat))):
            assert (_bool(ck, out[0]), R.decrypt_radix(ck, out[1:])) == (int(pos >= 0), max(pos, 0)), (hay, pat)
    for a, b in ((b"abcdefg", b"abcdefg"), (b"abcdefg", b"abcdefh"), (b"bbcdefg", b"abcdefz")):
        ea, eb = R.encrypt_string(ck, a), R.encrypt_string(ck, b)
        for out in both(lambda c: MG.sharded_eq(c, params, ea, eb, len(a))):
            assert _bool(ck, out) == int(a == b), (a, b)
        for op, w in (("lt", a < b), ("le", a <= b), ("gt", a > b), ("ge", a >= b)):
            for out in both(lambda c: MG.sharded_compare(c, params, op, ea, eb, len(a))):
                assert _bool(ck, out) == int(w), (op, a, b)
    s = b"Hello Zama, how is it going?"
    es = R.encrypt_string(ck, s)
    parts = both(lambda c: MG.sharded_case(c, params, "to_uppercase", es, len(s), gather=False))
    assert b"".join(R.decrypt_string(ck, x) for x in parts) == s.upper()
    # the exchange primitives on raw rows: sum and gather of two known blocks
    L = engines[0].p.big_len
    rows = [np.full((3, L), 7 + r, dtype=np.uint64) * np.arange(1, L + 1, dtype=np.uint64) for r in range(2)]
    summed = both(lambda c: c.to_host(c.all_reduce(torch.from_numpy(rows[c.rank].view(np.int64)).cuda())))
    gathered = both(lambda c: c.to_host(c.all_gather(torch.from_numpy(rows[c.rank].view(np.int64)).cuda()).reshape(-1, L)))
    for r in range(2):
        assert np.array_equal(summed[r], rows[0] + rows[1])
        assert np.array_equal(gathered[r], np.concatenate(rows))
    for c in comms:
        c.close()
    for e in engines:
        e.close()
