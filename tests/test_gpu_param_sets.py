"""GPU parity for the classic parameter sets outside N = 2048 / k = 1 / one PBS level: pbs_generic.cu and the
keyswitch with 2 ... 6 levels, against the oracle on identical seeded keys and ciphertexts.  Per set: keyswitch bit-exact; LUT rotation +
sample extraction bit-exact; one CMUX within the FFT tolerance of DESIGN section 4 (2^44 u64 torus units; the multi-level, k > 1 external
product exercises ggsw.rs:477-598 in full); KS-PBS decrypts to the LUT value for every message, with the phase error stated next to the
oracle's.  PARAM_MESSAGE_2_CARRY_2 itself runs through the generic kernel too (TFHE_B200_PBS_KERNEL=generic) and must agree with the tuned
kernels."""
import numpy as np
import pytest

from helpers import engine_params, oracle_partial_pbs

pytestmark = pytest.mark.gpu

SETS = ["1_0", "1_1", "2_0", "1_2", "1_3", "2_3", "1_4", "3_3", "4_3", "4_4"]   # N = 256 ... 32768 (4_4: the oracle's key generation alone
                                                                                # takes about a minute on 16 cores)


def _phase_error(ck, p, cts, want):
    delta = 2**63 // (p.msg_mod * p.carry_mod)
    errs = []
    for ct, v in zip(cts, want):
        d = (ck.decrypt_raw(ct) - int(v) * delta) % 2**64
        errs.append(min(d, 2**64 - d))
    return np.array(errs, dtype=np.float64)


@pytest.mark.parametrize("name", SETS)
def test_parameter_set(orc, name):
    import fhe_string_bounty_b200 as F
    p = orc.params(name)
    ck = orc.ClientKey(p, 0xB200 + 40)
    sk = orc.ServerKey(ck, 0xB300 + 40)
    space = p.msg_mod * p.carry_mod
    fs = [lambda x: x, lambda x: (3 * x + 1) % space, lambda x: int(x >= space // 2)]
    luts = np.stack([sk.generate_lookup_table(f)[0] for f in fs])
    eng = F.Engine(engine_params(p))
    eng.upload_ksk(sk.ksk)
    eng.upload_bsk_std(sk.bsk)
    eng.upload_luts(luts)

    n_cts = 96 if p.poly_size <= 8192 else 24
    vals = np.array([(v * 7) % space for v in range(space) for _ in fs][:n_cts])
    idx = np.array([i for _ in range(space) for i in range(len(fs))][:n_cts], dtype=np.uint32)
    cts = ck.encrypt_batch(vals)
    # keyswitch: exact integer arithmetic, including decomposer edge words
    edge = cts[:3].copy()
    edge[0, :] = np.uint64(2**64 - 1)
    edge[1, :] = np.random.default_rng(7).integers(0, 2**64, size=edge.shape[1], dtype=np.uint64)
    ks_in = np.concatenate([edge, cts[:6]])
    assert np.array_equal(eng.keyswitch_batch(ks_in), np.stack([sk.keyswitch(c) for c in ks_in])), name
    # blind rotation, step by step
    small = np.stack([sk.keyswitch(c) for c in cts[:3]])
    worst = 0
    for n_iters in (0, 1, 2):
        got = eng.pbs_batch(small, idx[:3], n_iters=n_iters)
        for b in range(3):
            want = oracle_partial_pbs(orc, sk, small[b], luts[idx[b]], n_iters)
            d = int(np.abs((got[b] - want).view(np.int64)).max())
            if n_iters == 0:
                assert d == 0, (name, "LUT rotation / sample extraction must be bit-exact")
            elif n_iters == 1:
                worst = max(worst, d)
                assert d <= 2**44, (name, b, np.log2(max(d, 1)))
    # full KS-PBS
    out = eng.ks_pbs_batch(cts, idx)
    want = np.array([fs[i](int(v)) for v, i in zip(vals, idx)])
    assert np.array_equal(ck.decrypt_batch(out), want), name
    assert np.array_equal(out, eng.ks_pbs_batch(cts, idx)), "deterministic"
    ref = sk.ks_pbs_batch(cts[:12], luts, idx[:12])
    assert np.array_equal(ck.decrypt_batch(ref), want[:12])
    e_gpu, e_cpu = _phase_error(ck, p, out, want), _phase_error(ck, p, ref, want[:12])
    margin = 2**63 // space // 2
    print(f"{name}: N={p.poly_size} k={p.glwe_dim} l={p.pbs_level}: one CMUX max|delta| 2^{np.log2(max(worst, 1)):.1f}; phase error gpu max "
          f"2^{np.log2(e_gpu.max()):.1f} rms 2^{np.log2(np.sqrt((e_gpu**2).mean())):.1f}, oracle max 2^{np.log2(e_cpu.max()):.1f} "
          f"(decoding margin 2^{np.log2(margin):.0f})")
    assert e_gpu.max() < margin / 4
    eng.close()


def test_generic_kernel_agrees_with_tuned_kernel(orc, keys_2_2, monkeypatch):
    """PARAM_MESSAGE_2_CARRY_2 through pbs_generic.cu: same words as pbs_v8.cu after LUT rotation, within 2^44 after one CMUX, same decrypted
    values after the full blind rotation."""
    import fhe_string_bounty_b200 as F
    p, ck, sk = keys_2_2
    acc, _ = sk.generate_lookup_table(lambda x: (5 * x + 2) % 16)

    def make():
        e = F.Engine(engine_params(p))
        e.upload_ksk(sk.ksk)
        e.upload_bsk_std(sk.bsk)
        e.upload_luts(acc[None, :])
        return e

    tuned = make()
    monkeypatch.setenv("TFHE_B200_PBS_KERNEL", "generic")
    gen = make()
    monkeypatch.delenv("TFHE_B200_PBS_KERNEL")
    vals = np.arange(32) % 16
    cts = ck.encrypt_batch(vals)
    small = tuned.keyswitch_batch(cts)
    assert np.array_equal(small, gen.keyswitch_batch(cts))
    assert np.array_equal(tuned.pbs_batch(small, None, n_iters=0), gen.pbs_batch(small, None, n_iters=0))
    d = np.abs((tuned.pbs_batch(small, None, n_iters=1) - gen.pbs_batch(small, None, n_iters=1)).view(np.int64)).max()
    assert d <= 2**44, np.log2(float(d))
    want = [(5 * int(v) + 2) % 16 for v in vals]
    assert list(ck.decrypt_batch(gen.ks_pbs_batch(cts, None))) == want
    assert list(ck.decrypt_batch(tuned.ks_pbs_batch(cts, None))) == want
    tuned.close()
    gen.close()


MULTI_BIT_SETS = ["multibit_1_1_g2", "multibit_2_2_g2", "multibit_3_3_g2", "multibit_1_1_g3", "multibit_3_3_g3"]


@pytest.mark.parametrize("name", MULTI_BIT_SETS)
def test_multi_bit_parameter_set(orc, name):
    """The other multi-bit sets of shortint/parameters/multi_bit.rs (grouping factor 2 and 3, N = 512 ... 8192, k up to 3, two PBS levels)
    on the generic kernel: keyswitch bit-exact, LUT rotation bit-exact, decrypted LUT values for every message, deterministic, phase error
    next to the oracle's deterministic multi-bit restatement."""
    import fhe_string_bounty_b200 as F
    p = orc.params(name)
    ck = orc.ClientKey(p, 0xB200 + 41)
    sk = orc.ServerKey(ck, 0xB300 + 41)
    space = p.msg_mod * p.carry_mod
    fs = [lambda x: x, lambda x: (3 * x + 1) % space, lambda x: int(x >= space // 2)]
    luts = np.stack([sk.generate_lookup_table(f)[0] for f in fs])
    eng = F.Engine(engine_params(p))
    eng.upload_ksk(sk.ksk)
    eng.upload_bsk_std(sk.bsk)
    eng.upload_luts(luts)
    vals = np.array([(v * 7) % space for v in range(space) for _ in fs][:96])
    idx = np.array([i for _ in range(space) for i in range(len(fs))][:96], dtype=np.uint32)
    cts = ck.encrypt_batch(vals)
    assert np.array_equal(eng.keyswitch_batch(cts[:6]), np.stack([sk.keyswitch(c) for c in cts[:6]])), name
    small = np.stack([sk.keyswitch(c) for c in cts[:3]])
    got0 = eng.pbs_batch(small, idx[:3], n_iters=0)
    for b in range(3):
        assert np.array_equal(got0[b], oracle_partial_pbs(orc, sk, small[b], luts[idx[b]], 0)), name
    out = eng.ks_pbs_batch(cts, idx)
    want = np.array([fs[i](int(v)) for v, i in zip(vals, idx)])
    assert np.array_equal(ck.decrypt_batch(out), want), name
    assert np.array_equal(out, eng.ks_pbs_batch(cts, idx)), "deterministic"
    ref = sk.ks_pbs_batch(cts[:12], luts, idx[:12])
    assert np.array_equal(ck.decrypt_batch(ref), want[:12])
    e_gpu, e_cpu = _phase_error(ck, p, out, want), _phase_error(ck, p, ref, want[:12])
    margin = 2**63 // space // 2
    print(f"{name}: N={p.poly_size} k={p.glwe_dim} l={p.pbs_level} g={p.grouping_factor}: phase error gpu max 2^{np.log2(e_gpu.max()):.1f} rms "
          f"2^{np.log2(np.sqrt((e_gpu**2).mean())):.1f}, oracle max 2^{np.log2(e_cpu.max()):.1f} (decoding margin 2^{np.log2(margin):.0f})")
    assert e_gpu.max() < margin / 4
    eng.close()


def test_generic_multi_bit_agrees_with_tuned_kernel(orc, keys_multibit, monkeypatch):
    """PARAM_MULTI_BIT_MESSAGE_2_CARRY_2_GROUP_3 through the generic kernel against pbs_multibit_v8.cu: identical LUT rotation, one group step
    within the FFT tolerance, identical decrypted values."""
    import fhe_string_bounty_b200 as F
    p, ck, sk = keys_multibit
    acc, _ = sk.generate_lookup_table(lambda x: (11 * x + 5) % 16)

    def make():
        e = F.Engine(engine_params(p))
        e.upload_ksk(sk.ksk)
        e.upload_bsk_std(sk.bsk)
        e.upload_luts(acc[None, :])
        return e

    tuned = make()
    monkeypatch.setenv("TFHE_B200_PBS_KERNEL", "generic")
    gen = make()
    monkeypatch.delenv("TFHE_B200_PBS_KERNEL")
    vals = np.arange(32) % 16
    cts = ck.encrypt_batch(vals)
    small = tuned.keyswitch_batch(cts)
    assert np.array_equal(tuned.pbs_batch(small, None, n_iters=0), gen.pbs_batch(small, None, n_iters=0))
    d = np.abs((tuned.pbs_batch(small, None, n_iters=1) - gen.pbs_batch(small, None, n_iters=1)).view(np.int64)).max()
    assert d <= 2**44, np.log2(float(d))
    want = [(11 * int(v) + 5) % 16 for v in vals]
    assert list(ck.decrypt_batch(gen.ks_pbs_batch(cts, None))) == want
    tuned.close()
    gen.close()


@pytest.mark.parametrize("name", ["1_3", "3_3"])
def test_string_ops_on_other_message_moduli_gpu(orc, name):
    """String programs on parameter sets whose message modulus is not 4: PARAM_MESSAGE_1_CARRY_3 (8 one-bit blocks per char, runs on
    the tuned N = 2048 kernels) and PARAM_MESSAGE_3_CARRY_3 (3 blocks per char, generic kernel, N = 8192); decrypted results against
    Python's bytes semantics."""
    import fhe_string_bounty_b200 as F
    from fhe_string_bounty_b200.host import Program
    from oracle import radix as R
    from helpers import engine_params
    p = orc.params(name)
    ck = orc.ClientKey(p, 0xB230)
    sk = orc.ServerKey(ck, 0xB231)
    eng = F.Engine(engine_params(p))
    eng.upload_ksk(sk.ksk)
    eng.upload_bsk_std(sk.bsk)
    cases = [(b"Zama", b"Zama"), (b"Zama", b"Zamb"), (b"abcab", b"ca")] if name == "3_3" else \
        [(b"hello", b"hello"), (b"hello", b"hellp"), (b"abd", b"abc"), (b"abcabd", b"abd"), (b"zz~", b"zz")]
    for a, b in cases:
        ins = np.concatenate([R.encrypt_string(ck, a), R.encrypt_string(ck, b)])
        want = {"eq": a == b, "ne": a != b, "lt": a < b, "ge": a >= b, "contains": b in a, "starts_with": a.startswith(b)}
        for op, w in want.items():
            P = Program("string_" + op, (len(a), len(b)), params=engine_params(p))
            out = P.run(eng, ins)
            assert ck.decrypt_message_and_carry(out[0]) == int(w), (name, op, a, b)
    eng.close()


def test_tuned_n512_kernel(orc):
    """PARAM_MESSAGE_1_CARRY_1_KS_PBS (N = 512, k = 3) on pbs_n512.cu -- 16 x 16 FFT in half-warps, eight ciphertexts per SM sharing one
    TMA key ring -- for its 1-, 4- and 8-ciphertext instances: LUT rotation + sample extraction bit-exact against the oracle, one and two
    CMUXes within 2^44 (max stated), every message through three LUTs decrypts, phase error bounded, identical words on re-run, and
    agreement with the generic kernel on the same inputs."""
    import torch
    import fhe_string_bounty_b200 as F
    p = orc.params("1_1")
    ck = orc.ClientKey(p, 0xB200 + 41)
    sk = orc.ServerKey(ck, 0xB300 + 41)
    space = p.msg_mod * p.carry_mod
    fs = [lambda x: x, lambda x: (3 * x + 1) % space, lambda x: int(x >= space // 2)]
    luts = np.stack([sk.generate_lookup_table(f)[0] for f in fs])
    eng = F.Engine(engine_params(p))
    eng.upload_ksk(sk.ksk)
    eng.upload_bsk_std(sk.bsk)
    eng.upload_luts(luts)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    base_vals = np.array([v for v in range(space) for _ in fs])
    base_idx = np.array([i for _ in range(space) for i in range(len(fs))], dtype=np.uint32)
    base = ck.encrypt_batch(base_vals)
    worst = 0
    for batch in (3, sms + 5, 8 * sms + 5):                     # -> pbs_n512_kernel<1>, <4>, <8>
        reps = -(-batch // len(base))
        cts = np.tile(base, (reps, 1))[:batch]
        vals, idx = np.tile(base_vals, reps)[:batch], np.tile(base_idx, reps)[:batch]
        small = eng.keyswitch_batch(cts)
        eng.set_tuning("tuned512_min", 1)
        part = [eng.pbs_batch(small, idx, n_iters=k) for k in (0, 1, 2)]
        out = eng.ks_pbs_batch(cts, idx)
        assert np.array_equal(out, eng.ks_pbs_batch(cts, idx)), "deterministic"
        eng.set_tuning("tuned512_min", 1 << 30)                 # generic kernel
        gen = [eng.pbs_batch(small, idx, n_iters=k) for k in (0, 1)]
        assert np.array_equal(part[0], gen[0])
        assert np.abs((part[1] - gen[1]).view(np.int64)).max() <= 2**44
        for b in (0, 1, batch - 1):
            for k in (0, 1, 2):
                want = oracle_partial_pbs(orc, sk, small[b], luts[idx[b]], k)
                d = int(np.abs((part[k][b] - want).view(np.int64)).max())
                if k == 0:
                    assert d == 0, (batch, b, "LUT rotation / sample extraction must be bit-exact")
                else:
                    worst = max(worst, d)
                    assert d <= 2**44, (batch, b, k, np.log2(max(d, 1)))
        want = np.array([fs[i](int(v)) for v, i in zip(vals, idx)])
        assert np.array_equal(ck.decrypt_batch(out), want), batch
        err = _phase_error(ck, p, out[:64], want[:64])
        assert err.max() < 2**63 // space // 8
    print(f"pbs_n512: one / two CMUX max|delta| 2^{np.log2(max(worst, 1)):.1f}")
    eng.close()


def test_tuned_n8192_kernel(orc):
    """PARAM_MESSAGE_3_CARRY_3_KS_PBS (N = 8192, k = 1, two levels) on pbs_n8192.cu -- one ciphertext per two-SM cluster, 16 x 16 x 16 FFT,
    spectra swapped per level with st.async -- against the oracle (LUT rotation bit-exact, one CMUX within 2^44 with the
    maximum stated) and against the generic kernel; every message through three LUTs decrypts; identical words on re-run; a batch wider
    than the number of clusters that fit the GPU."""
    import torch
    import fhe_string_bounty_b200 as F
    p = orc.params("3_3")
    ck = orc.ClientKey(p, 0xB200 + 42)
    sk = orc.ServerKey(ck, 0xB300 + 42)
    space = p.msg_mod * p.carry_mod
    fs = [lambda x: x, lambda x: (3 * x + 1) % space, lambda x: int(x >= space // 2)]
    luts = np.stack([sk.generate_lookup_table(f)[0] for f in fs])
    eng = F.Engine(engine_params(p))
    eng.upload_ksk(sk.ksk)
    eng.upload_bsk_std(sk.bsk)
    eng.upload_luts(luts)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    base_vals = np.array([(v * 5) % space for v in range(48)])
    base_idx = (np.arange(48) % 3).astype(np.uint32)
    base = ck.encrypt_batch(base_vals)
    worst = 0
    for batch in (3, sms // 2 + 7):
        reps = -(-batch // len(base))
        cts = np.tile(base, (reps, 1))[:batch]
        vals, idx = np.tile(base_vals, reps)[:batch], np.tile(base_idx, reps)[:batch]
        small = eng.keyswitch_batch(cts)
        eng.set_tuning("tuned8192", 1)
        part = [eng.pbs_batch(small, idx, n_iters=k) for k in (0, 1)]
        out = eng.ks_pbs_batch(cts, idx)
        assert np.array_equal(out, eng.ks_pbs_batch(cts, idx)), "deterministic"
        eng.set_tuning("tuned8192", 0)                          # generic kernel
        gen = [eng.pbs_batch(small, idx, n_iters=k) for k in (0, 1)]
        assert np.array_equal(part[0], gen[0])
        assert np.abs((part[1] - gen[1]).view(np.int64)).max() <= 2**44
        for b in (0, 1, batch - 1):
            for k in (0, 1):     # (after two CMUXes a 2^33 rounding difference flips level-2 digits, boundaries 2^34 apart: DESIGN section 4)
                want = oracle_partial_pbs(orc, sk, small[b], luts[idx[b]], k)
                d = int(np.abs((part[k][b] - want).view(np.int64)).max())
                if k == 0:
                    assert d == 0, (batch, b, "LUT rotation / sample extraction must be bit-exact")
                else:
                    worst = max(worst, d)
                    assert d <= 2**44, (batch, b, k, np.log2(max(d, 1)))
        want = np.array([fs[i](int(v)) for v, i in zip(vals, idx)])
        assert np.array_equal(ck.decrypt_batch(out), want), batch
        err = _phase_error(ck, p, out[:48], want[:48])
        assert err.max() < 2**63 // space // 8
    print(f"pbs_n8192: one CMUX max|delta| 2^{np.log2(max(worst, 1)):.1f}")
    eng.close()
