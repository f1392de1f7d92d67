"""Shared helpers for the parity tests (oracle side)."""
import ctypes as C

import numpy as np


def oracle_partial_pbs(orc, sk, lwe_small, lut, n_iters, log2_q=64):
    """bootstrap.rs:254-316 restated with the oracle primitives, stopping after n_iters mask elements;
    returns the sample-extracted LWE (same shape as a full PBS output)."""
    L = orc.lib()
    p = sk.p
    N, k1 = p.poly_size, p.glwe_dim + 1
    lg = int(N).bit_length() - 1
    b_hat = L.orc_modulus_switch(int(lwe_small[p.lwe_dim]), lg)
    acc = np.zeros(k1 * N, dtype=np.uint64)
    for q in range(k1):
        tmp = np.zeros(N, dtype=np.uint64)
        L.orc_monomial_div(tmp, np.ascontiguousarray(lut[q * N:(q + 1) * N]), N, b_hat)
        acc[q * N:(q + 1) * N] = tmp
    for i in range(n_iters):
        if int(lwe_small[i]) == 0:
            continue
        a_hat = L.orc_modulus_switch(int(lwe_small[i]), lg)
        ct1 = np.zeros_like(acc)
        for q in range(k1):
            tmp = np.zeros(N, dtype=np.uint64)
            L.orc_monomial_mul_and_subtract(tmp, np.ascontiguousarray(acc[q * N:(q + 1) * N]), N, a_hat)
            ct1[q * N:(q + 1) * N] = tmp
        L.orc_add_external_product_f64(C.byref(p), sk.fourier, i, acc, ct1)
    if log2_q < 64:   # bootstrap.rs:318-330: SignedDecomposer(log2_q, 1).closest_representable on every coefficient, then extract
        shift = np.uint64(64 - log2_q - 1)
        acc = (((acc >> shift) + np.uint64(1)) & ~np.uint64(1)) << shift
    out = np.zeros(p.big_dim + 1, dtype=np.uint64)
    L.orc_sample_extract0(C.byref(p), np.ascontiguousarray(acc), out)
    return out


def engine_params(orc_params):
    return dict(lwe_dim=orc_params.lwe_dim, glwe_dim=orc_params.glwe_dim, poly_size=orc_params.poly_size,
                pbs_base_log=orc_params.pbs_base_log, pbs_level=orc_params.pbs_level,
                ks_base_log=orc_params.ks_base_log, ks_level=orc_params.ks_level,
                grouping_factor=orc_params.grouping_factor, msg_mod=orc_params.msg_mod, carry_mod=orc_params.carry_mod)


def phase_error(ck, cts, expected_values):
    """|decrypted phase - expected plaintext| in u64 torus units, per ciphertext."""
    errs = []
    for ct, v in zip(cts, expected_values):
        ph = ck.decrypt_raw(ct)
        d = (ph - (int(v) << 59)) % 2**64
        errs.append(min(d, 2**64 - d))
    return np.array(errs, dtype=np.float64)


def simulate_program_clear(ir, inputs_clear, total_mod):
    """Noise-free execution of a recorded program on PLAINTEXT values: a slot holds its value modulo 2 * total_mod (message space plus
    the padding bit), leveled instructions are exact, a PBS is the table lookup the blind rotation performs -- f(x) below the padding
    bit, -f(x - total_mod) above it (negacyclic LUT, shortint/server_key/mod.rs:763-781).  Any operand that reaches the padding bit is
    reported: returns (outputs, n_padding_hits).  Used for programs far too large for the CPU oracle's real PBS."""
    M = 2 * total_mod
    delta = (1 << 63) // total_mod
    val = np.zeros(ir.n_slots, dtype=np.int64)
    val[: ir.n_inputs] = np.asarray(inputs_clear, dtype=np.int64)
    hits = 0
    for lv in range(len(ir.level_lin_off) - 1):
        for li in range(int(ir.level_lin_off[lv]), int(ir.level_lin_off[lv + 1])):
            out, tb, te = (int(v) for v in ir.lin[li])
            body = int(ir.lin_body[li])
            assert body % delta == 0, "plaintext bodies are multiples of delta"
            acc = body // delta
            for t in range(tb, te):
                acc += int(ir.term_coef[t]) * int(val[ir.term_slot[t]])
            val[out] = acc % M
        for j in range(int(ir.level_pbs_off[lv]), int(ir.level_pbs_off[lv + 1])):
            src, dst, lut = (int(v) for v in ir.pbs[j])
            x = int(val[src])
            if x >= total_mod:
                hits += 1
                val[dst] = (-int(ir.lut_tables[lut][x - total_mod])) % M
            else:
                val[dst] = int(ir.lut_tables[lut][x])
    return val[ir.outputs], hits


def string_blocks_clear(s: bytes):
    """4 little-endian 2-bit blocks per char"""
    return [(ch >> (2 * b)) & 3 for ch in s for b in range(4)]
