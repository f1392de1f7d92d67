"""Host-layer logic on CPU (no GPU): the product's C++ scheduler records integer / string operations as level-batched
programs; here every program is (a) checked for the level shapes SURVEY.md Appendix B derives from the reference's
trees, and (b) executed with the ORACLE's CPU KS-PBS on a toy parameter set and compared with clear-text semantics and
with the oracle's own block-by-block restatement of the reference trees (oracle/radix.py)."""
import numpy as np
import pytest

from fhe_string_bounty_b200.host import Program
from helpers import engine_params


def test_level_shapes_match_reference_trees():
    assert Program("string_eq", (8, 8)).level_widths == [32, 3, 1]                         # config 1
    assert Program("string_ne", (8, 8)).level_widths == [32, 3, 1]
    P = Program("string_contains", (256, 16))                                             # config 3
    assert P.level_widths == [15424, 1205, 241, 17, 2, 1] and P.n_pbs == 16890
    assert Program("string_lt", (128, 128)).level_widths == [256, 128, 64, 32, 16, 8, 4, 2, 1, 1]  # config 5
    assert Program("string_to_lowercase", (1024,)).n_pbs == 4 * 1024                       # fused case circuit
    assert Program("radix_eq", (4,)).level_widths == [4, 1]                                # FheUint8 eq
    assert Program("radix_lt", (128,)).level_widths == [64, 32, 16, 8, 4, 2, 1, 1]         # FheUint256 compare
    # empty / degenerate operands fold to trivial ciphertexts on the host (scalar_comparison.rs:151-153)
    assert Program("string_eq", (0, 0)).n_pbs == 0
    assert Program("string_eq", (3, 4)).n_pbs == 0
    assert Program("string_contains", (4, 0)).n_pbs == 0


def test_accumulators_match_oracle_generate_lookup_table(orc, toy_keys, keys_2_2):
    """fill_accumulator restated in the product (engine/mod.rs:72-128) == the oracle's, bit for bit."""
    for p, ck, sk in (toy_keys, keys_2_2):
        P = Program("string_find", (6, 2), params=engine_params(p))
        ir = P.ir()
        accs = P.accumulators()
        for t, deg, acc in zip(ir.lut_tables, ir.lut_degrees, accs):
            want, wdeg = sk.generate_lookup_table(lambda x, t=t: int(t[x]))
            assert np.array_equal(acc, want) and int(deg) == wdeg


def _run(orc, keys, op, args, inputs, clear=None):
    from oracle import radix as R
    p, ck, sk = keys
    P = Program(op, args, clear=clear, params=engine_params(p))
    return R.run_program(P.ir(), sk, inputs), P


def _dec_bool(ck, ct):
    return ck.decrypt_message_and_carry(ct)


STR_CASES = [
    (b"hello", b"hello"), (b"hello", b"hellp"), (b"Hello", b"hello"), (b"abc", b"abd"), (b"abd", b"abc"),
    (b"", b""), (b"a", b""), (b"ab", b"abc"), (b"abc", b"ab"), (b"zzz", b"zz~"), (b"\x00\x7f", b"\x00\x7f"),
]


@pytest.mark.parametrize("a,b", STR_CASES)
def test_string_comparisons_toy(orc, toy_keys, a, b):
    from oracle import radix as R
    p, ck, sk = toy_keys
    ins = np.concatenate([R.encrypt_string(ck, a), R.encrypt_string(ck, b)]) if (a or b) else np.zeros((0, p.big_dim + 1), dtype=np.uint64)
    want = {"eq": a == b, "ne": a != b, "lt": a < b, "le": a <= b, "gt": a > b, "ge": a >= b,
            "starts_with": a.startswith(b), "ends_with": a.endswith(b), "contains": b in a,
            "eq_ignore_case": a.lower() == b.lower()}
    for op, w in want.items():
        out, P = _run(orc, toy_keys, "string_" + op, (len(a), len(b)), ins)
        assert _dec_bool(ck, out[0]) == int(w), (op, a, b)
        assert tuple(P.ir().output_degree_noise[0]) in ((1, 1), (1, 0), (0, 0)), "boolean block, nominal noise or trivial"


def test_string_ops_against_reference_style_oracle(orc, toy_keys):
    """Same inputs through (i) the product's recorded program and (ii) the oracle's block-by-block restatement of
    unchecked_eq / unchecked_ne / unchecked_lt..ge: decrypted booleans agree; PBS counts agree with the reference's."""
    from oracle import radix as R
    p, ck, sk = toy_keys
    rng = np.random.default_rng(21)
    for trial in range(6):
        n = int(rng.integers(1, 7))
        a = bytes(rng.integers(0x20, 0x7F, size=n).tolist())
        b = a if trial % 2 == 0 else bytes(rng.integers(0x20, 0x7F, size=n).tolist())   # forced-equal case, tests_cases_comparisons.rs:43-48
        ea, eb = R.encrypt_string(ck, a), R.encrypt_string(ck, b)
        key = R.ShortintServerKey(sk)
        isk = R.IntegerServerKey(key)
        la, lb = list(ea), list(eb)
        # lexicographic order = big-endian integer: chars reversed
        ra = [blk for i in reversed(range(n)) for blk in ea[4 * i:4 * i + 4]]
        rb = [blk for i in reversed(range(n)) for blk in eb[4 * i:4 * i + 4]]
        ref = {"eq": isk.unchecked_eq(la, lb), "ne": isk.unchecked_ne(la, lb), "lt": isk.unchecked_lt(ra, rb),
               "le": isk.unchecked_le(ra, rb), "gt": isk.unchecked_gt(ra, rb), "ge": isk.unchecked_ge(ra, rb)}
        clear = {"eq": a == b, "ne": a != b, "lt": a < b, "le": a <= b, "gt": a > b, "ge": a >= b}
        for op in ref:
            out, P = _run(orc, toy_keys, "string_" + op, (n, n), np.concatenate([ea, eb]))
            assert _dec_bool(ck, out[0]) == _dec_bool(ck, ref[op]) == int(clear[op]), (op, a, b)
        # PBS count of eq: B blocks + count tree (comparison.rs:10-33)
        P = Program("string_eq", (n, n), params=engine_params(p))
        B = 4 * n
        tree, m = 0, B
        while m > 1:
            m = -(-m // 15)
            tree += m
        assert P.n_pbs == B + tree


def test_case_conversion_tutorial_kat(orc, toy_keys):
    """docs/tutorials/ascii_fhe_string.md:140-153: 'Hello Zama, how is it going?' -> upper / lower."""
    from oracle import radix as R
    p, ck, sk = toy_keys
    s = b"Hello Zama, how is it going?"
    enc = R.encrypt_string(ck, s)
    out, _ = _run(orc, toy_keys, "string_to_uppercase", (len(s),), enc)
    assert R.decrypt_string(ck, out) == b"HELLO ZAMA, HOW IS IT GOING?"
    out, _ = _run(orc, toy_keys, "string_to_lowercase", (len(s),), enc)
    assert R.decrypt_string(ck, out) == b"hello zama, how is it going?"
    edge = bytes([0x40, 0x41, 0x5A, 0x5B, 0x60, 0x61, 0x7A, 0x7B, 0x00, 0x7F, 0x30, 0x20])
    enc = R.encrypt_string(ck, edge)
    out, _ = _run(orc, toy_keys, "string_to_lowercase", (len(edge),), enc)
    assert R.decrypt_string(ck, out) == edge.lower()
    out, _ = _run(orc, toy_keys, "string_to_uppercase", (len(edge),), enc)
    assert R.decrypt_string(ck, out) == edge.upper()


def test_find_and_contains_toy(orc, toy_keys):
    from oracle import radix as R
    p, ck, sk = toy_keys
    for hay, pat in [(b"abcabcab", b"cab"), (b"abcabcab", b"abc"), (b"abcabcab", b"zzz"), (b"aaaa", b"aa"), (b"xyz", b"xyz"), (b"xy", b"xyz")]:
        ins = np.concatenate([R.encrypt_string(ck, hay), R.encrypt_string(ck, pat)])
        out, P = _run(orc, toy_keys, "string_find", (len(hay), len(pat)), ins)
        found, idx = _dec_bool(ck, out[0]), R.decrypt_radix(ck, out[1:])
        want = hay.find(pat)
        assert found == int(want >= 0) and idx == max(want, 0), (hay, pat, found, idx)
        # rfind: the last match (same circuit over the windows in descending order), reference-shaped and packed
        for op in ("string_rfind", "string_rfind_packed"):
            out, _ = _run(orc, toy_keys, op, (len(hay), len(pat)), ins)
            want_r = hay.rfind(pat)
            assert (_dec_bool(ck, out[0]), R.decrypt_radix(ck, out[1:])) == (int(want_r >= 0), max(want_r, 0)), (op, hay, pat)
        # clear pattern variant: the pattern is a trivial string, equality against it still costs the same PBS
        out, _ = _run(orc, toy_keys, "string_contains", (len(hay),), R.encrypt_string(ck, hay), clear=pat.decode())
        assert _dec_bool(ck, out[0]) == int(pat in hay)


def test_radix_ops_toy(orc, toy_keys):
    """integer tests_cases_comparisons.rs style: random operands + forced-equal case, unchecked flavour; add and cmux."""
    from oracle import radix as R
    p, ck, sk = toy_keys
    rng = np.random.default_rng(33)
    nb = 6
    for trial in range(5):
        x = int(rng.integers(0, 4**nb))
        y = x if trial == 0 else int(rng.integers(0, 4**nb))
        ins = np.stack(R.encrypt_radix(ck, x, nb) + R.encrypt_radix(ck, y, nb))
        for op, w in {"eq": x == y, "ne": x != y, "lt": x < y, "le": x <= y, "gt": x > y, "ge": x >= y}.items():
            out, _ = _run(orc, toy_keys, "radix_" + op, (nb,), ins)
            assert _dec_bool(ck, out[0]) == int(w), (op, x, y)
        out, _ = _run(orc, toy_keys, "radix_add", (nb,), ins)
        assert R.decrypt_radix(ck, out) == (x + y) % 4**nb
        for s in (y, 7, 4**nb + 5):
            xin = np.stack(R.encrypt_radix(ck, x, nb))
            for op, w in {"scalar_eq": x == s, "scalar_lt": x < s, "scalar_gt": x > s}.items():
                out, _ = _run(orc, toy_keys, "radix_" + op, (nb, s), xin)
                assert _dec_bool(ck, out[0]) == int(w), (op, x, s)
        for cond in (0, 1):
            cin = np.concatenate([ck.encrypt(cond)[None, :], ins])
            out, _ = _run(orc, toy_keys, "radix_if_then_else", (nb,), cin)
            assert R.decrypt_radix(ck, out) == (x if cond else y)


def test_shortint_level_toy(orc, toy_keys):
    """shortint.rs:366-462 restated through the program API: univariate LUT on every value incl. carries, bivariate (2xy)%4."""
    from oracle import radix as R
    p, ck, sk = toy_keys
    f = [(3 * v + 1) % 16 for v in range(16)]
    ins = ck.encrypt_batch(range(16))
    out, _ = _run(orc, toy_keys, "shortint_apply_lut", [16] + f, ins)
    assert [ck.decrypt_message_and_carry(c) for c in out] == f
    g = [(2 * x * y) % 4 for x in range(4) for y in range(4)]
    xs, ys = np.repeat(np.arange(4), 4), np.tile(np.arange(4), 4)
    ins = np.stack([ck.encrypt(int(v)) for v in xs] + [ck.encrypt(int(v)) for v in ys])
    out, _ = _run(orc, toy_keys, "shortint_bivariate_lut", [16] + g, ins)
    assert [ck.decrypt_message_and_carry(c) for c in out] == g


def test_unknown_op_and_bad_args_fail_loudly():
    from fhe_string_bounty_b200 import NativeError
    with pytest.raises(NativeError):
        Program("string_frobnicate", (3, 3))
    with pytest.raises(NativeError):
        Program("radix_eq", ())


def test_many_independent_pairs_share_levels(orc, toy_keys):
    """throughput mode: N independent eq's in one program = the single-op tree widened N times, same results per pair"""
    from oracle import radix as R
    p, ck, sk = toy_keys
    P1 = Program("string_eq", (3, 3), params=engine_params(p))
    PN = Program("string_eq_many", (3, 3, 5), params=engine_params(p))
    assert PN.level_widths == [5 * w for w in P1.level_widths]
    pairs = [(b"abc", b"abc"), (b"abc", b"abd"), (b"zzz", b"zzz"), (b"a b", b"a_b"), (b"xyz", b"xyz")]
    ins = np.concatenate([np.concatenate([R.encrypt_string(ck, a), R.encrypt_string(ck, b)]) for a, b in pairs])
    out = R.run_program(PN.ir(), sk, ins)
    assert [ck.decrypt_message_and_carry(c) for c in out] == [int(a == b) for a, b in pairs]
    # the same with packed block equalities (one PBS per pair of blocks): fewer PBS, same booleans
    PP = Program("string_eq_many_packed", (3, 3, 5), params=engine_params(p))
    assert PP.n_pbs < PN.n_pbs and PP.level_widths[0] * 2 == PN.level_widths[0]
    out = R.run_program(PP.ir(), sk, ins)
    assert [ck.decrypt_message_and_carry(c) for c in out] == [int(a == b) for a, b in pairs]


def test_packed_equality_halves_the_pbs_and_agrees(orc, toy_keys):
    """*_packed ops: one PBS per pair of blocks (pack + lwe_sub + LUT[x == 0], comparator.rs:193-221 restated); results
    identical to the reference-shaped programs on the same inputs, including every 4-bit difference sign."""
    from oracle import radix as R
    p, ck, sk = toy_keys
    assert Program("string_eq_packed", (8, 8)).level_widths == [16, 2, 1]
    P = Program("string_contains_packed", (256, 16))
    assert P.level_widths == [7712, 723, 241, 17, 2, 1] and P.n_pbs == 8696
    # nibble pairs of both signs: equality through the padding-bit trick must be exact for negative differences too
    for x, y in [(0, 0), (0, 15), (15, 0), (7, 8), (8, 7), (9, 9), (15, 15), (1, 0), (0, 1), (12, 3)]:
        ins = np.stack(R.encrypt_radix(ck, x, 2) + R.encrypt_radix(ck, y, 2))
        out, _ = _run(orc, toy_keys, "radix_eq_packed", (2,), ins)
        assert _dec_bool(ck, out[0]) == int(x == y), (x, y)
    for a, b in STR_CASES + [(b"abcabcab", b"cab"), (b"abcabcab", b"cba")]:
        ins = np.concatenate([R.encrypt_string(ck, a), R.encrypt_string(ck, b)]) if (a or b) else np.zeros((0, p.big_dim + 1), dtype=np.uint64)
        for op, w in {"eq": a == b, "contains": b in a, "starts_with": a.startswith(b), "ends_with": a.endswith(b)}.items():
            out, _ = _run(orc, toy_keys, f"string_{op}_packed", (len(a), len(b)), ins)
            assert _dec_bool(ck, out[0]) == int(w), (op, a, b)
    hay, pat = b"abcabcab", b"cab"
    ins = np.concatenate([R.encrypt_string(ck, hay), R.encrypt_string(ck, pat)])
    out, _ = _run(orc, toy_keys, "string_find_packed", (len(hay), len(pat)), ins)
    assert _dec_bool(ck, out[0]) == 1 and R.decrypt_radix(ck, out[1:]) == hay.find(pat)


def test_find_full_size_first_match_positions(orc, toy_keys):
    """find at BASELINE config-3 size (256/16, 241 windows = 18 blocks of 14, i.e. past the 15-block carry group) on the toy
    set: first match in the first block, in the middle, in the last block, repeated matches, and no match."""
    from oracle import radix as R
    p, ck, sk = toy_keys
    rng = np.random.default_rng(99)
    P = Program("string_find_packed", (256, 16), params=engine_params(p))
    ir = P.ir()
    assert len(P.level_widths) == 9      # the first-match flag is folded into the index-digit selection (one level fewer)
    base = bytes(rng.integers(ord("a"), ord("z") + 1, size=256).tolist())
    pat = b"QRSTUVWXYZ012345"
    for positions in ([0], [5], [13, 14], [120], [209, 230], [215], [240], [3, 100, 240], []):
        hay = bytearray(base)
        for pos in positions:
            hay[pos:pos + 16] = pat
        hay = bytes(hay)
        ins = np.concatenate([R.encrypt_string(ck, hay), R.encrypt_string(ck, pat)])
        out = R.run_program(ir, sk, ins)
        want = hay.find(pat)
        assert ck.decrypt_message_and_carry(out[0]) == int(want >= 0), positions
        assert R.decrypt_radix(ck, out[1:]) == max(want, 0), positions


@pytest.mark.parametrize("hay,pat", [
    (b"ab" * 300, b"ab"),                    # 599 windows, a match in every block of 14: the carry of the block-prefix sums is live
    (b"b" * 1024, b"b"),                     # 1024 windows, every window matches
    (b"a" * 700 + b"b", b"ab"),              # single match in the last group
    (b"xy" * 256 + b"ab" * 256, b"ab"),      # first match exactly at a group boundary region
    (b"a" * 900, b"b"),                      # no match at all
])
def test_find_long_haystack_dense_matches(hay, pat):
    """More than 30 blocks of 14 windows (> 420 windows): the cleaned carry of the 'some earlier block matched' prefix sums must never
    push a PBS operand past total_mod - 1 (round-1 bug: carry + 15 flags = 16 reached the padding bit and the carry was lost, so a
    later window was reported as first).  Executed noise-free on plaintexts (the program is far too large for the CPU oracle)."""
    from helpers import simulate_program_clear, string_blocks_clear
    for op, want_pos in (("string_find", hay.find(pat)), ("string_rfind", hay.rfind(pat))):
        P = Program(op, (len(hay), len(pat)))
        out, hits = simulate_program_clear(P.ir(), string_blocks_clear(hay) + string_blocks_clear(pat), 16)
        assert hits == 0, f"{op}: {hits} PBS operands reached the padding bit"
        found, idx = int(out[0]), sum(int(d) << (2 * i) for i, d in enumerate(out[1:]))
        assert (found, idx) == (int(want_pos >= 0), max(want_pos, 0)), (op, found, idx, want_pos)


def test_pbs_rejects_operand_past_message_space():
    """Program::pbs refuses a ciphertext whose degree exceeds total_mod - 1 instead of recording a wrong LUT evaluation"""
    from fhe_string_bounty_b200 import NativeError
    with pytest.raises(NativeError, match="degree"):
        Program("bool_sum_finish", (16, 0))
    Program("bool_sum_finish", (15, 1))


@pytest.mark.parametrize("degree", [3, 6, 9, 15])
def test_default_comparisons_on_dirty_carries(degree):
    """Operands with non-empty carries (tests_cases_comparisons.rs:81-97 raises the degree with unchecked_add before comparing): the
    default forms propagate first (message / carry extraction, shifted add, single-carry propagation) and then run the unchecked
    circuit.  Noise-free execution on plaintext blocks holding up to `degree`; the compared VALUE is sum_i block_i 4^i mod 4^B."""
    from helpers import simulate_program_clear
    rng = np.random.default_rng(degree)
    nb = 8
    for trial in range(12):
        x = rng.integers(0, degree + 1, size=nb)
        y = x.copy() if trial % 3 == 0 else rng.integers(0, degree + 1, size=nb)
        vx = sum(int(v) << (2 * i) for i, v in enumerate(x)) % 4**nb
        vy = sum(int(v) << (2 * i) for i, v in enumerate(y)) % 4**nb
        out, _ = simulate_program_clear(Program("radix_full_propagate", (nb, degree)).ir(), list(x), 16)
        assert all(int(d) < 4 for d in out) and sum(int(d) << (2 * i) for i, d in enumerate(out)) == vx
        for op, w in (("eq", vx == vy), ("ne", vx != vy), ("lt", vx < vy), ("le", vx <= vy), ("gt", vx > vy), ("ge", vx >= vy)):
            out, _ = simulate_program_clear(Program("radix_default_" + op, (nb, degree)).ir(), list(x) + list(y), 16)
            assert int(out[0]) == int(w), (op, degree, list(x), list(y))
    # fresh operands take the no-propagation branch (comparison.rs:213-214): same program as the unchecked form
    assert Program("radix_default_lt", (nb, 3)).n_pbs == Program("radix_lt", (nb,)).n_pbs
    assert Program("radix_default_lt", (nb, 6)).n_pbs > Program("radix_lt", (nb,)).n_pbs


@pytest.mark.parametrize("msg_mod,carry_mod", [(2, 8), (8, 8), (16, 16)])
def test_string_ops_on_other_message_moduli(orc, msg_mod, carry_mod):
    """A char is an FheUint8 = ceil(8 / log2(message_modulus)) blocks (integer/encryption.rs:69-83): 8 one-bit blocks for the MESSAGE_1
    sets, 3 for MESSAGE_3, 2 for MESSAGE_4.  eq / ne / the comparator (needs >= 4 bits of message + carry, comparator.rs:51-60) /
    contains / starts_with / ends_with are composed from the radix methods and do not care; the fused case conversion, find and the
    null-padded model are written for 2-bit blocks and say so."""
    from oracle import radix as R
    p = orc.params("toy")
    p.msg_mod, p.carry_mod = msg_mod, carry_mod
    p.poly_size = max(256, 16 * msg_mod * carry_mod)    # a box of the lookup table must stay wider than the modulus-switch rounding
    ck = orc.ClientKey(p, 0xB210 + msg_mod)
    sk = orc.ServerKey(ck, 0xB220 + msg_mod)
    keys = (p, ck, sk)
    bpc = R.blocks_per_char(msg_mod)
    assert bpc == {2: 8, 8: 3, 16: 2}[msg_mod]
    assert R.decrypt_string(ck, R.encrypt_string(ck, b"Az~\x00\x7f")) == b"Az~\x00\x7f"
    for a, b in [(b"hello", b"hello"), (b"hello", b"hellp"), (b"abd", b"abc"), (b"ab", b"abc"), (b"zz~", b"zz"), (b"abcabd", b"abd")]:
        ins = np.concatenate([R.encrypt_string(ck, a), R.encrypt_string(ck, b)])
        want = {"eq": a == b, "ne": a != b, "lt": a < b, "le": a <= b, "gt": a > b, "ge": a >= b,
                "starts_with": a.startswith(b), "ends_with": a.endswith(b), "contains": b in a}
        for op, w in want.items():
            out, P = _run(orc, keys, "string_" + op, (len(a), len(b)), ins)
            assert P.n_inputs == bpc * (len(a) + len(b))
            assert _dec_bool(ck, out[0]) == int(w), (msg_mod, op, a, b)
    for op, args in (("string_to_lowercase", (3,)), ("string_find", (4, 2)), ("pstring_len", (4,))):
        with pytest.raises(Exception, match="message modulus 4"):
            Program(op, args, params=engine_params(p))
