"""Null-padded FheString operations (secret length; fhe_string_bounty_b200/csrc/host/padded.h) on CPU: every operation is recorded by
the product's C++ scheduler and executed (a) noise-free on plaintexts -- exhaustive-ish over edge cases: empty strings, strings that
fill the capacity, whitespace-only strings, patterns longer than the haystack -- against Rust's str semantics (len, is_empty, ==, <,
trim_start / trim_end / trim with ASCII char::is_whitespace, strip_prefix / strip_suffix, +, repeat, contains / starts_with /
ends_with), and (b) with the CPU oracle's real KS-PBS on the toy parameter set for a few cases (degrees and noise as recorded)."""
import numpy as np
import pytest

from fhe_string_bounty_b200.host import Program
from helpers import engine_params, simulate_program_clear, string_blocks_clear


def pad(s: bytes, cap: int) -> bytes:
    return s + b"\0" * (cap - len(s))


def blocks_to_bytes(blocks) -> bytes:
    return bytes(sum(int(d) << (2 * i) for i, d in enumerate(blocks[k:k + 4])) for k in range(0, len(blocks), 4))


def run_clear(op, args, ins, clear=None):
    P = Program(op, args, clear=clear)
    out, hits = simulate_program_clear(P.ir(), ins, 16)
    # only the comparator reaches the padding bit, and on purpose (comparator.rs:193-221)
    assert hits == 0 or op.split("_")[-1] in ("lt", "le", "gt", "ge"), (op, hits)
    return out, P


UNARY = [b"", b"a", b"hello", b"  hi  ", b"\t\n x y \r ", b"            ", b"abcdefghijkl", b"   ", b"ab   ", b"\x0b\x0cz", b" \x1f "]


@pytest.mark.parametrize("s", UNARY)
def test_len_trim_strip_clear_simulation(s):
    cap = 12
    ins = string_blocks_clear(pad(s, cap))
    out, _ = run_clear("pstring_len", (cap,), ins)
    assert sum(int(d) << (2 * i) for i, d in enumerate(out)) == len(s)
    out, _ = run_clear("pstring_is_empty", (cap,), ins)
    assert int(out[0]) == int(len(s) == 0)
    ws = b" \t\n\r\x0b\x0c"                     # char::is_whitespace restricted to ASCII
    for op, want in (("trim_end", s.rstrip(ws)), ("trim_start", s.lstrip(ws)), ("trim", s.strip(ws))):
        out, _ = run_clear("pstring_" + op, (cap,), ins)
        got = blocks_to_bytes(out)
        assert got == pad(want, cap), (op, s, got)          # the result is itself a well-formed padded string
    for pat in (b"ab", b"he", b" ", b"hello", b"kl", b"l", b"abcdefghijklm"):
        out, _ = run_clear("pstring_strip_prefix", (cap,), ins, clear=pat)
        want = s[len(pat):] if s.startswith(pat) else s
        assert int(out[0]) == int(s.startswith(pat)) and blocks_to_bytes(out[1:]) == pad(want, cap), (s, pat)
        out, _ = run_clear("pstring_strip_suffix", (cap,), ins, clear=pat)
        want = s[:-len(pat)] if s.endswith(pat) else s
        assert int(out[0]) == int(s.endswith(pat)) and blocks_to_bytes(out[1:]) == pad(want, cap), (s, pat)


BINARY = [(b"abababab", b"ab"), (b"abababab", b""), (b"aaaa", b"aa"), (b"xaxbxc", b"xc"),
          (b"hello", b"hello"), (b"hello", b"hell"), (b"", b""), (b"abc", b"abd"), (b"b", b"abc"), (b"abcdefgh", b"efgh"),
          (b"abcab", b"ab"), (b"abcab", b""), (b"", b"x"), (b"xyz", b"yz"), (b"xyzxyz", b"zx"), (b"abcdefgh", b"abcdef"), (b"aab", b"ab")]


@pytest.mark.parametrize("a,b", BINARY)
def test_padded_binary_ops_clear_simulation(a, b):
    ca, cb = 8, 6
    ins = string_blocks_clear(pad(a, ca)) + string_blocks_clear(pad(b, cb))
    want = {"eq": a == b, "ne": a != b, "lt": a < b, "le": a <= b, "gt": a > b, "ge": a >= b, "contains": b in a,
            "starts_with": a.startswith(b), "ends_with": a.endswith(b)}
    for op, w in want.items():
        out, _ = run_clear("pstring_" + op, (ca, cb), ins)
        assert int(out[0]) == int(w), (op, a, b)
    out, _ = run_clear("pstring_concat", (ca, cb), ins)
    assert blocks_to_bytes(out) == pad(a + b, ca + cb), (a, b)
    # str::find / str::rfind: (found, byte index); the empty pattern is found at 0 / at len
    for op, pos in (("find", a.find(b)), ("rfind", a.rfind(b))):
        out, _ = run_clear("pstring_" + op, (ca, cb), ins)
        got = (int(out[0]), sum(int(d) << (2 * i) for i, d in enumerate(out[1:])))
        assert got == (int(pos >= 0), max(pos, 0)), (op, a, b, got)


def test_repeat_and_shapes():
    for s in (b"ab", b"", b"xyz"):
        out, _ = run_clear("pstring_repeat", (3, 3), string_blocks_clear(pad(s, 3)))
        assert blocks_to_bytes(out) == pad(s * 3, 9)
    # wide levels: what the engine is built for (a 32-char trim is ~2 k KS-PBS in ~20 levels)
    P = Program("pstring_trim", (32,))
    assert P.n_pbs > 1500 and max(P.level_widths) >= 128
    assert Program("pstring_len", (255,)).n_outputs == 4           # 255 < 4^4
    from fhe_string_bounty_b200 import NativeError
    with pytest.raises(NativeError, match="clear pattern"):
        Program("pstring_strip_prefix", (8,))


def test_padded_ops_with_real_pbs_toy(orc, toy_keys):
    """the recorded programs executed with the oracle's CPU KS-PBS on the toy parameter set (degree / noise bookkeeping is real)"""
    from oracle import radix as R
    p, ck, sk = toy_keys
    cap = 6
    for s in (b" ab ", b"abc"):
        enc = R.encrypt_string(ck, pad(s, cap))
        out = R.run_program(Program("pstring_len", (cap,), params=engine_params(p)).ir(), sk, enc)
        assert R.decrypt_radix(ck, out) == len(s)
        out = R.run_program(Program("pstring_trim", (cap,), params=engine_params(p)).ir(), sk, enc)
        assert R.decrypt_string(ck, out) == pad(s.strip(), cap)
        out = R.run_program(Program("pstring_strip_suffix", (cap,), clear="b ", params=engine_params(p)).ir(), sk, enc)
        want = s[:-2] if s.endswith(b"b ") else s
        assert ck.decrypt_message_and_carry(out[0]) == int(s.endswith(b"b ")) and R.decrypt_string(ck, out[1:]) == pad(want, cap)
    a, b = b"ab", b"c"
    enc = np.concatenate([R.encrypt_string(ck, pad(a, 3)), R.encrypt_string(ck, pad(b, 2))])
    out = R.run_program(Program("pstring_concat", (3, 2), params=engine_params(p)).ir(), sk, enc)
    assert R.decrypt_string(ck, out) == pad(a + b, 5)
    out = R.run_program(Program("pstring_ends_with", (3, 2), params=engine_params(p)).ir(), sk, enc)
    assert ck.decrypt_message_and_carry(out[0]) == 0
