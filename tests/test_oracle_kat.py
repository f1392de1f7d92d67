"""Pins the CPU oracle against every known-answer / property the reference's own tests hold for the
KS-PBS path.  CPU only."""
import ctypes as C

import numpy as np
import pytest


def test_decomposer_doc_kats(orc):
    L = orc.lib()
    # commons/math/decomposition/decomposer.rs:94-95
    assert L.orc_closest_representable_u32(1_340_987_234, 4, 3) == 1_341_128_704
    # decomposer.rs:142: decompose(1).count() == 3
    assert len(orc.decompose(1, 4, 3)) == 3


def test_decomposer_properties(orc):
    """decomposition/tests.rs:32-136: terms in [-B/2, B/2], recompose == closest_representable,
    rounding is idempotent; plus order = level l first (iter.rs:45-50)."""
    L = orc.lib()
    rng = np.random.default_rng(1)
    for base_log, level in [(3, 5), (23, 1), (7, 2), (21, 1), (4, 3), (8, 4)]:
        for x in rng.integers(0, 2**64, size=300, dtype=np.uint64):
            x = int(x)
            digits = orc.decompose(x, base_log, level)
            half = 1 << (base_log - 1)
            assert all(-half <= d <= half for d in digits)
            closest = L.orc_closest_representable(x, base_log, level)
            rec = 0
            for idx, d in enumerate(digits):
                lvl = level - idx  # digits[0] is level `level`
                rec += d << (64 - base_log * lvl)
            assert rec % 2**64 == closest
            assert L.orc_closest_representable(closest, base_log, level) == closest
            err = (x - closest) % 2**64
            err = min(err, 2**64 - err)
            assert err <= 1 << (64 - base_log * level - 1)


def test_modulus_switch(orc):
    L = orc.lib()
    # fft_impl/common.rs:26-43: round(x / 2^(64-log2(2N))) in [0, 2N]
    for x in [0, 1, 2**52 - 1, 2**51, 2**51 - 1, 2**63, 2**64 - 1, 2**64 - 2**51, 2**64 - 2**51 - 1]:
        want = (x + 2**51) >> 52
        assert L.orc_modulus_switch(x, 11) == want
    assert L.orc_modulus_switch(2**64 - 1, 11) == 4096  # the documented "may return 2N" case


def test_monomial_doc_kats(orc):
    L = orc.lib()
    # polynomial_algorithms.rs:305-313 (u8 doc test), restated mod 2^64
    inp = np.array([1, 2, 3], dtype=np.uint64)
    out = np.zeros(3, dtype=np.uint64)
    L.orc_monomial_div(out, inp, 3, 2)
    assert [int(v) for v in out] == [3, 2**64 - 1, 2**64 - 2]
    # polynomial_algorithms.rs:944-982: mul then div is the identity; mul_and_subtract = mul - id
    rng = np.random.default_rng(2)
    for N in (8, 256, 2048):
        x = rng.integers(0, 2**64, size=N, dtype=np.uint64)
        for deg in [0, 1, N - 1, N, N + 3, 2 * N - 1, 2 * N, int(rng.integers(0, 2 * N))]:
            a = np.zeros(N, dtype=np.uint64); b = np.zeros(N, dtype=np.uint64); c = np.zeros(N, dtype=np.uint64)
            L.orc_monomial_mul(a, x, N, deg)
            L.orc_monomial_div(b, a, N, deg)
            assert np.array_equal(b, x)
            L.orc_monomial_mul_and_subtract(c, x, N, deg)
            assert np.array_equal(c, a - x)


def _schoolbook_negacyclic(a, b):
    """exact product in Z[X]/(X^N+1) with python ints"""
    N = len(a)
    out = [0] * N
    for i, ai in enumerate(a):
        if ai == 0:
            continue
        for j, bj in enumerate(b):
            if i + j < N:
                out[i + j] += ai * bj
            else:
                out[i + j - N] -= ai * bj
    return out


def test_fft_roundtrip_u64(orc):
    """fft64/math/fft/tests.rs:9-80: forward_as_torus then backward: |delta| < 2^14."""
    L = orc.lib()
    rng = np.random.default_rng(3)
    for N in (256, 1024, 2048, 4096):
        poly = rng.integers(0, 2**64, size=N, dtype=np.uint64)
        re = np.zeros(N // 2); im = np.zeros(N // 2)
        L.orc_fft_forward_torus(N, poly, re, im)
        back = np.zeros(N, dtype=np.uint64)
        L.orc_fft_add_backward_torus(N, back, re, im)
        d = (back - poly).view(np.int64)
        assert np.abs(d).max() < 2**14


def test_fft_product_vs_schoolbook(orc):
    """fft64/math/fft/tests.rs:82-222: torus poly * 16-bit integer poly vs exact schoolbook,
    |delta| <= 2^(64 - (52 - 16 - log2 N))."""
    L = orc.lib()
    rng = np.random.default_rng(4)
    for N in (256, 2048):
        M = N // 2
        a = rng.integers(0, 2**64, size=N, dtype=np.uint64)
        b = rng.integers(-2**15, 2**15, size=N, dtype=np.int64)
        ar = np.zeros(M); ai = np.zeros(M); br = np.zeros(M); bi = np.zeros(M)
        L.orc_fft_forward_torus(N, a, ar, ai)
        L.orc_fft_forward_integer(N, b.view(np.uint64), br, bi)
        pr = ar * br - ai * bi
        pi = ar * bi + ai * br
        got = np.zeros(N, dtype=np.uint64)
        L.orc_fft_add_backward_torus(N, got, pr, pi)
        exact = _schoolbook_negacyclic([int(v) for v in a], [int(v) for v in b])
        exact = np.array([v % 2**64 for v in exact], dtype=np.uint64)
        d = (got - exact).view(np.int64)
        log2N = N.bit_length() - 1
        assert np.abs(d).max() <= 2 ** (64 - (52 - 16 - log2N))


def test_lut_layout_and_trivial_pbs(orc, toy_keys):
    """engine/mod.rs:94-127 layout; server_key/mod.rs:763-781 trivial PBS incl. padding-bit negation."""
    p, ck, sk = toy_keys
    L = orc.lib()
    f = lambda x: (3 * x + 1) % 16
    acc, degree = sk.generate_lookup_table(f)
    N, box = p.poly_size, p.poly_size // 16
    delta = 2**59
    assert degree == max(f(i) for i in range(16))
    assert not acc[:N].any()
    body = acc[N:]
    for i in range(16):
        # after rotate-left by box/2, box i occupies [i*box - box/2, i*box + box/2)
        assert int(body[i * box]) == f(i) * delta
    assert int(body[N - 1]) == (-f(0) * delta) % 2**64  # negated first half-box wrapped to the end
    for v in range(32):
        got = L.orc_trivial_pbs(C.byref(p), v * delta, acc)
        want = f(v) * delta if v < 16 else (-f(v % 16) * delta) % 2**64
        assert got == want


def test_keyswitch_decrypts(orc, toy_keys):
    """algorithms/test/lwe_keyswitch.rs:8-109: encrypt -> KS -> decrypt == msg for every message."""
    p, ck, sk = toy_keys
    for m in range(16):
        ct = ck.encrypt_with_carry(m)
        ks = sk.keyswitch(ct)
        assert orc.lib().orc_decode(C.byref(p), ck.decrypt_small_raw(ks)) == m


def test_pbs_identity_all_messages_toy(orc, toy_keys):
    """algorithms/test/lwe_programmable_bootstrapping.rs:69-166 (identity LUT, all messages) for
    both oracle flavours (f64 FFT and exact integer)."""
    p, ck, sk = toy_keys
    ident, _ = sk.generate_lookup_table(lambda x: x)
    for m in range(16):
        ks = sk.keyswitch(ck.encrypt_with_carry(m))
        assert ck.decrypt_message_and_carry(sk.pbs(ks, ident)) == m
        assert ck.decrypt_message_and_carry(sk.pbs(ks, ident, exact=True)) == m


def test_external_product_f64_vs_exact(orc, keys_2_2):
    """One CMUX step at the real parameter set: the f64 flavour must sit within the reference's own
    FFT tolerance (fft/tests.rs:166-167 scaled to 23-bit digits) of the exact integer product."""
    p, ck, sk = keys_2_2
    L = orc.lib()
    rng = np.random.default_rng(5)
    glwe = rng.integers(0, 2**64, size=p.lut_len, dtype=np.uint64)
    out_f = np.zeros(p.lut_len, dtype=np.uint64)
    out_e = np.zeros(p.lut_len, dtype=np.uint64)
    L.orc_add_external_product_f64(C.byref(p), sk.fourier, 3, out_f, glwe)
    ggsw_len = p.pbs_level * 4 * p.poly_size
    L.orc_add_external_product_exact(C.byref(p), sk.bsk[3 * ggsw_len:4 * ggsw_len].copy(), out_e, glwe)
    d = np.abs((out_f - out_e).view(np.int64)).max()
    assert d <= 2 ** (64 - (52 - 23 - 11)), f"max|f64-exact| = 2^{np.log2(d):.1f}"


def test_ks_pbs_2_2_all_messages_and_bivariate(orc, keys_2_2):
    """shortint/server_key/tests/shortint.rs:366-462: KS-PBS with identity and a bivariate
    (2*x*y)%4 LUT on PARAM_MESSAGE_2_CARRY_2_KS_PBS; trivial PBS agrees after decryption
    (shortint.rs:3233-3296)."""
    p, ck, sk = keys_2_2
    L = orc.lib()
    ident, _ = sk.generate_lookup_table(lambda x: x)
    biv, _ = sk.generate_lookup_table_bivariate(lambda x, y: (2 * x * y) % 4)
    luts = np.stack([ident, biv])
    cts = ck.encrypt_batch(range(16))
    idx = np.zeros(16, dtype=np.uint32)
    out = sk.ks_pbs_batch(cts, luts, idx)
    assert list(ck.decrypt_batch(out)) == list(range(16))
    idx[:] = 1
    out = sk.ks_pbs_batch(cts, luts, idx)
    want = [(2 * (v // 4) * (v % 4)) % 4 for v in range(16)]
    assert list(ck.decrypt_batch(out)) == want
    for v in range(16):
        triv = L.orc_trivial_pbs(C.byref(p), v * 2**59, biv)
        assert L.orc_decode(C.byref(p), triv) == want[v]


def test_batch_equals_sequential(orc, keys_2_2):
    """lwe_keyswitch.rs:93 (par == seq bit-exact) restated for the OpenMP batch entry point."""
    p, ck, sk = keys_2_2
    ident, _ = sk.generate_lookup_table(lambda x: x)
    cts = ck.encrypt_batch([1, 7, 12])
    out_b, ks_b = sk.ks_pbs_batch(cts, ident, want_ks=True)
    for i in range(3):
        ks = sk.keyswitch(cts[i])
        assert np.array_equal(ks, ks_b[i])
        assert np.array_equal(sk.pbs(ks, ident), out_b[i])


def test_power_of_two_modulus_pbs_oracle(orc, toy_keys):
    """bootstrap.rs:318-330 (the reference runs its PBS tests with q = 2^63 as well): outputs are multiples of 2^(64 - log2 q) and still
    decrypt to the LUT value; with q = 2^64 the function is the plain PBS."""
    p, ck, sk = toy_keys
    acc, _ = sk.generate_lookup_table(lambda x: (x + 3) % 16)
    for v in (0, 7, 15):
        small = sk.keyswitch(ck.encrypt_with_carry(v))
        assert np.array_equal(sk.pbs_pow2_modulus(small, acc, 64), sk.pbs(small, acc))
        for log2_q in (63, 32, 20):
            out = sk.pbs_pow2_modulus(small, acc, log2_q)
            assert not np.any(out & np.uint64((1 << (64 - log2_q)) - 1))
            assert ck.decrypt_message_and_carry(out) == (v + 3) % 16
