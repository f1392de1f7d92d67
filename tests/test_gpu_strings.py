"""GPU tests of the host layer at PARAM_MESSAGE_2_CARRY_2_KS_PBS: every string / radix operation recorded by the C++
scheduler is executed on the B200 (one leveled + one keyswitch + one PBS launch per tree level) and its DECRYPTED result
must be bit-exact against clear-text semantics and against the same program executed with the CPU oracle."""
import numpy as np
import pytest

from fhe_string_bounty_b200.host import Program
from helpers import engine_params

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng(keys_2_2):
    import fhe_string_bounty_b200 as F
    p, ck, sk = keys_2_2
    e = F.Engine(engine_params(p))
    e.upload_ksk(sk.ksk)
    e.upload_bsk_std(sk.bsk)
    yield e
    e.close()


def _bool(ck, ct):
    return ck.decrypt_message_and_carry(ct)


def test_eq_8_chars_config1(orc, keys_2_2, eng):
    """BASELINE config 1: eq of two 8-char ASCII strings; equal with prob 1/2 (forced copy); GPU == oracle == clear."""
    from oracle import radix as R
    p, ck, sk = keys_2_2
    rng = np.random.default_rng(0xB200 + 0)
    P = Program("string_eq", (8, 8), params=engine_params(p))
    Pn = Program("string_ne", (8, 8), params=engine_params(p))
    assert P.level_widths == [32, 3, 1]
    for trial in range(4):
        a = bytes(rng.integers(0x20, 0x7F, size=8).tolist())
        b = a if trial % 2 == 0 else bytes(rng.integers(0x20, 0x7F, size=8).tolist())
        ins = np.concatenate([R.encrypt_string(ck, a), R.encrypt_string(ck, b)])
        got = P.run(eng, ins)
        assert _bool(ck, got[0]) == int(a == b)
        assert _bool(ck, Pn.run(eng, ins)[0]) == int(a != b)
        if trial < 2:
            ref = R.run_program(P.ir(), sk, ins)
            assert _bool(ck, ref[0]) == int(a == b)


def test_string_ops_small(orc, keys_2_2, eng):
    from oracle import radix as R
    p, ck, sk = keys_2_2
    cases = [(b"needle in hay", b"in h"), (b"needle in hay", b"hay"), (b"needle in hay", b"hey"), (b"abc", b"abd"),
             (b"abd", b"abc"), (b"same", b"same"), (b"ab", b"abc"), (b"MiXeD", b"mixed")]
    for a, b in cases:
        ins = np.concatenate([R.encrypt_string(ck, a), R.encrypt_string(ck, b)])
        want = {"eq": a == b, "lt": a < b, "le": a <= b, "gt": a > b, "ge": a >= b, "contains": b in a,
                "starts_with": a.startswith(b), "ends_with": a.endswith(b), "eq_ignore_case": a.lower() == b.lower()}
        for op, w in want.items():
            P = Program("string_" + op, (len(a), len(b)), params=engine_params(p))
            out = P.run(eng, ins)
            assert _bool(ck, out[0]) == int(w), (op, a, b)
        P = Program("string_find", (len(a), len(b)), params=engine_params(p))
        out = P.run(eng, ins)
        f = a.find(b)
        assert _bool(ck, out[0]) == int(f >= 0) and R.decrypt_radix(ck, out[1:]) == max(f, 0), (a, b)
        out = Program("string_rfind", (len(a), len(b)), params=engine_params(p)).run(eng, ins)
        f = a.rfind(b)
        assert _bool(ck, out[0]) == int(f >= 0) and R.decrypt_radix(ck, out[1:]) == max(f, 0), ("rfind", a, b)


def test_case_conversion_kat_gpu(orc, keys_2_2, eng):
    """docs/tutorials/ascii_fhe_string.md:140-153"""
    from oracle import radix as R
    p, ck, sk = keys_2_2
    s = b"Hello Zama, how is it going?"
    enc = R.encrypt_string(ck, s)
    up = Program("string_to_uppercase", (len(s),), params=engine_params(p)).run(eng, enc)
    lo = Program("string_to_lowercase", (len(s),), params=engine_params(p)).run(eng, enc)
    assert R.decrypt_string(ck, up) == b"HELLO ZAMA, HOW IS IT GOING?"
    assert R.decrypt_string(ck, lo) == b"hello zama, how is it going?"


def test_radix_ops_gpu(orc, keys_2_2, eng):
    from oracle import radix as R
    p, ck, sk = keys_2_2
    rng = np.random.default_rng(5)
    nb = 16   # 32-bit operands
    for trial in range(3):
        x = int(rng.integers(0, 2**32))
        y = x if trial == 0 else int(rng.integers(0, 2**32))
        ins = np.stack(R.encrypt_radix(ck, x, nb) + R.encrypt_radix(ck, y, nb))
        for op, w in {"eq": x == y, "ne": x != y, "lt": x < y, "le": x <= y, "gt": x > y, "ge": x >= y}.items():
            assert _bool(ck, Program("radix_" + op, (nb,), params=engine_params(p)).run(eng, ins)[0]) == int(w), (op, x, y)
        out = Program("radix_add", (nb,), params=engine_params(p)).run(eng, ins)
        assert R.decrypt_radix(ck, out) == (x + y) % 2**32
        cin = np.concatenate([ck.encrypt(trial % 2)[None, :], ins])
        out = Program("radix_if_then_else", (nb,), params=engine_params(p)).run(eng, cin)
        assert R.decrypt_radix(ck, out) == (x if trial % 2 else y)


def test_contains_256_16_config3(orc, keys_2_2, eng):
    """BASELINE config 3 at full size: 256-char haystack, 16-char pattern, 16 890 PBS in 6 levels; pattern copied from a
    random offset with prob 1/2.  Decrypted result vs clear text (size-independent property: contains(h, h[o:o+16]) == 1)."""
    from oracle import radix as R
    p, ck, sk = keys_2_2
    rng = np.random.default_rng(0xB200 + 3)
    P = Program("string_contains", (256, 16), params=engine_params(p))
    F = Program("string_find", (256, 16), params=engine_params(p))
    for trial in range(2):
        hay = bytes(rng.integers(ord("a"), ord("z") + 1, size=256).tolist())
        off = int(rng.integers(0, 241))
        pat = hay[off:off + 16] if trial == 0 else bytes(rng.integers(ord("a"), ord("z") + 1, size=16).tolist())
        ins = np.concatenate([R.encrypt_string(ck, hay), R.encrypt_string(ck, pat)])
        out = P.run(eng, ins)
        assert _bool(ck, out[0]) == int(pat in hay)
        out = F.run(eng, ins)
        f = hay.find(pat)
        assert _bool(ck, out[0]) == int(f >= 0) and R.decrypt_radix(ck, out[1:]) == max(f, 0)
    print(f"contains 256/16: {P.n_pbs} PBS, {P.last_ms():.1f} ms on device; find: {F.n_pbs} PBS, {F.last_ms():.1f} ms")


def test_packed_equality_ops_gpu(orc, keys_2_2, eng):
    """packed-equality variants at the real parameter set (their PBS input carries the noise of a packed subtraction, like
    the reference's comparisons): decrypted results vs clear text, full-size contains 256/16 with half the PBS."""
    from oracle import radix as R
    p, ck, sk = keys_2_2
    rng = np.random.default_rng(0xB200 + 33)
    for a, b in [(b"needle in hay", b"in h"), (b"needle in hay", b"hey"), (b"same", b"same"), (b"samf", b"same"), (b"MiXeD", b"mixed")]:
        ins = np.concatenate([R.encrypt_string(ck, a), R.encrypt_string(ck, b)])
        for op, w in {"eq": a == b, "contains": b in a, "starts_with": a.startswith(b), "ends_with": a.endswith(b),
                      "eq_ignore_case": a.lower() == b.lower()}.items():
            out = Program(f"string_{op}_packed", (len(a), len(b)), params=engine_params(p)).run(eng, ins)
            assert _bool(ck, out[0]) == int(w), (op, a, b)
    # every ordered pair of 4-bit values, 256 independent 2-block equalities through one batched program
    xs = np.repeat(np.arange(16), 16)
    ys = np.tile(np.arange(16), 16)
    P = Program("string_contains_packed", (256, 16), params=engine_params(p))
    F = Program("string_find_packed", (256, 16), params=engine_params(p))
    assert P.n_pbs == 8696
    for trial in range(2):
        hay = bytes(rng.integers(ord("a"), ord("z") + 1, size=256).tolist())
        off = int(rng.integers(0, 241))
        pat = hay[off:off + 16] if trial == 0 else bytes(rng.integers(ord("a"), ord("z") + 1, size=16).tolist())
        ins = np.concatenate([R.encrypt_string(ck, hay), R.encrypt_string(ck, pat)])
        assert _bool(ck, P.run(eng, ins)[0]) == int(pat in hay)
        out = F.run(eng, ins)
        f = hay.find(pat)
        assert _bool(ck, out[0]) == int(f >= 0) and R.decrypt_radix(ck, out[1:]) == max(f, 0)
    print(f"contains 256/16 packed: {P.n_pbs} PBS, {P.last_ms():.1f} ms; find packed: {F.n_pbs} PBS, {F.last_ms():.1f} ms")
    for x, y in zip(xs[::7], ys[::7]):
        ins = np.stack(R.encrypt_radix(ck, int(x), 2) + R.encrypt_radix(ck, int(y), 2))
        assert _bool(ck, Program("radix_eq_packed", (2,), params=engine_params(p)).run(eng, ins)[0]) == int(x == y), (x, y)


def test_config2_batch_4096_real_ciphertexts(orc, keys_2_2, eng):
    """BASELINE config 2 at its smallest named size: 4096 independent LUT evaluations, 16 LUTs round-robin, real seeded keys
    and fresh ciphertexts; every decrypted output must equal LUT(message) (size-independent property), through both the
    chunked host-buffer entry point and the device entry point."""
    import torch
    p, ck, sk = keys_2_2
    fs = [lambda x, k=k: (x * (2 * k + 1) + k) % 16 for k in range(16)]
    luts = np.stack([sk.generate_lookup_table(f)[0] for f in fs])
    eng.upload_luts(luts)
    rng = np.random.default_rng(0xB200 + 2)
    vals = rng.integers(0, 16, size=4096)
    idx = (np.arange(4096) % 16).astype(np.uint32)
    cts = ck.encrypt_batch(vals)
    want = np.array([fs[i](int(v)) for v, i in zip(vals, idx)])
    out = eng.ks_pbs_batch(cts, idx)
    assert np.array_equal(ck.decrypt_batch(out), want)
    d_in = torch.from_numpy(cts.view(np.int64)).cuda()
    d_idx = torch.from_numpy(idx.view(np.int32)).cuda()
    d_out = torch.empty_like(d_in)
    s = torch.cuda.Stream()
    eng.ks_pbs_batch_device(d_in, d_idx, d_out, 4096, s.cuda_stream)
    s.synchronize()
    assert np.array_equal(d_out.cpu().numpy().view(np.uint64), out), "chunked host path and device path run the same kernels"


def test_config4_case_ops_1024_chars(orc, keys_2_2, eng):
    """BASELINE config 4: to_lowercase / to_uppercase / eq_ignore_case on 1024-char mixed-case strings."""
    from oracle import radix as R
    p, ck, sk = keys_2_2
    rng = np.random.default_rng(0xB200 + 4)
    alphabet = np.frombuffer(b"abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789 ", dtype=np.uint8)
    a = bytes(rng.choice(alphabet, size=1024).tolist())
    b = a.swapcase() if True else a
    ea = R.encrypt_string(ck, a)
    lo = Program("string_to_lowercase", (1024,), params=engine_params(p))
    up = Program("string_to_uppercase", (1024,), params=engine_params(p))
    assert R.decrypt_string(ck, lo.run(eng, ea)) == a.lower()
    assert R.decrypt_string(ck, up.run(eng, ea)) == a.upper()
    eb = R.encrypt_string(ck, b)
    eic = Program("string_eq_ignore_case", (1024, 1024), params=engine_params(p))
    assert _bool(ck, eic.run(eng, np.concatenate([ea, eb]))[0]) == 1
    c = bytearray(b)
    c[700] = ord("#")
    assert _bool(ck, eic.run(eng, np.concatenate([ea, R.encrypt_string(ck, bytes(c))]))[0]) == 0
    print(f"to_lowercase 1024 chars: {lo.n_pbs} PBS, {lo.last_ms():.1f} ms; eq_ignore_case: {eic.n_pbs} PBS, {eic.last_ms():.1f} ms")
