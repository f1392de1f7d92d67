"""GPU parity tests (run on the B200 box with -m gpu): the CUDA path through the C ABI vs the CPU oracle
on identical seeded keys and ciphertexts.

Bars (BASELINE.json north_star / SURVEY.md 8c):
  * keyswitch outputs: bit-exact,
  * one external product (one CMUX) from identical inputs: max |delta| <= 2^44 u64 torus units vs the
    oracle's f64 flavour (the reference's own FFT tolerance for 23-bit digits is 2^46, fft/tests.rs:166-167),
  * full PBS: decrypted values bit-exact; decrypted phase error far below the decoding margin 2^58.
Full-PBS ciphertext words are NOT comparable word by word: a 2^39 rounding difference in one external
product flips level-1 digits (granularity 2^41) of the next, which re-randomises the mask while leaving
the phase unchanged ("What parity means for PBS outputs")."""
import numpy as np
import pytest

from helpers import engine_params, oracle_partial_pbs, phase_error

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


def _luts(sk):
    fs = [lambda x: x, lambda x: x % 4, lambda x: x // 4, lambda x: int(x == 5), lambda x: (3 * x + 1) % 16]
    accs = [sk.generate_lookup_table(f)[0] for f in fs]
    return fs, np.stack(accs)


def test_keyswitch_bit_exact(orc, keys_2_2, eng):
    p, ck, sk = keys_2_2
    rng = np.random.default_rng(10)
    for batch in (1, 3, 70):
        cts = ck.encrypt_batch(rng.integers(0, 16, size=batch))
        if batch == 70:  # also arbitrary (non-ciphertext) words, incl. edge values of the decomposer
            cts[0, :] = 0
            cts[1, :] = np.uint64(2**64 - 1)
            cts[2, :] = rng.integers(0, 2**64, size=cts.shape[1], dtype=np.uint64)
            cts[3, :8] = np.array([2**48, 2**48 - 1, 2**48 + 1, 2**63, 2**63 - 2**48, 2**49, 3 * 2**48, 2**64 - 2**48], dtype=np.uint64)
        got = eng.keyswitch_batch(cts)
        want = np.stack([sk.keyswitch(c) for c in cts])
        assert np.array_equal(got, want), f"batch {batch}: {np.argwhere(got != want)[:5]}"


def test_one_cmux_matches_oracle(orc, keys_2_2, eng):
    p, ck, sk = keys_2_2
    fs, luts = _luts(sk)
    eng.upload_luts(luts)
    cts = ck.encrypt_batch([0, 5, 9, 15])
    small = np.stack([sk.keyswitch(c) for c in cts])
    idx = np.array([0, 3, 4, 1], dtype=np.uint32)
    worst = 0
    for n_iters in (0, 1):
        got = eng.pbs_batch(small, idx, n_iters=n_iters)
        for b in range(len(cts)):
            want = oracle_partial_pbs(orc, sk, small[b], luts[idx[b]], n_iters)
            d = np.abs((got[b] - want).view(np.int64)).max()
            worst = max(worst, int(d))
            if n_iters == 0:
                assert d == 0, "LUT rotation / sample extraction must be bit-exact"
            if n_iters == 1:
                assert d <= 2**44, f"n_iters={n_iters} ct {b}: max|delta| = 2^{np.log2(max(d, 1)):.1f}"
    print(f"max |delta| GPU vs oracle-f64 after one CMUX: 2^{np.log2(max(worst, 1)):.1f}")


def test_ks_pbs_all_messages_all_luts(orc, keys_2_2, eng):
    p, ck, sk = keys_2_2
    fs, luts = _luts(sk)
    eng.upload_luts(luts)
    vals = np.array([v for v in range(16) for _ in fs])
    idx = np.array([i for _ in range(16) for i in range(len(fs))], dtype=np.uint32)
    cts = ck.encrypt_batch(vals)
    out = eng.ks_pbs_batch(cts, idx)
    want = np.array([fs[i](int(v)) for v, i in zip(vals, idx)])
    got = ck.decrypt_batch(out)
    assert np.array_equal(got, want)
    err_gpu = phase_error(ck, out, want)
    ref = sk.ks_pbs_batch(cts, luts, idx)
    assert np.array_equal(ck.decrypt_batch(ref), want)
    err_cpu = phase_error(ck, ref, want)
    print(f"phase error (u64 torus units): gpu max 2^{np.log2(err_gpu.max()):.1f} rms 2^{np.log2(np.sqrt((err_gpu**2).mean())):.1f}; "
          f"oracle max 2^{np.log2(err_cpu.max()):.1f} rms 2^{np.log2(np.sqrt((err_cpu**2).mean())):.1f}")
    assert err_gpu.max() < 2**54           # decoding margin is 2^58
    assert np.sqrt((err_gpu**2).mean()) < 2 * np.sqrt((err_cpu**2).mean()) + 2**40


def test_bivariate_and_ragged_batches(orc, keys_2_2, eng):
    """shortint.rs:432-462 bivariate (2*x*y)%4, and batch sizes that do not fill a keyswitch tile."""
    p, ck, sk = keys_2_2
    biv, _ = sk.generate_lookup_table_bivariate(lambda x, y: (2 * x * y) % 4)
    eng.upload_luts(biv[None, :])
    for batch in (1, 2, 63, 65):
        vals = np.arange(batch) % 16
        out = eng.ks_pbs_batch(ck.encrypt_batch(vals), None)
        want = [(2 * (v // 4) * (v % 4)) % 4 for v in vals]
        assert list(ck.decrypt_batch(out)) == want
    assert eng.ks_pbs_batch(np.zeros((0, p.big_dim + 1), dtype=np.uint64)).shape[0] == 0


def test_device_entry_point_matches_host(orc, keys_2_2, eng):
    import torch
    p, ck, sk = keys_2_2
    fs, luts = _luts(sk)
    eng.upload_luts(luts)
    vals = np.arange(32) % 16
    cts = ck.encrypt_batch(vals)
    idx = (np.arange(32) % len(fs)).astype(np.uint32)
    host = eng.ks_pbs_batch(cts, idx)
    d_in = torch.from_numpy(cts.view(np.int64)).cuda()
    d_idx = torch.from_numpy(idx.view(np.int32)).cuda()
    d_out = torch.empty_like(d_in)
    eng.ks_pbs_batch_device(d_in, d_idx, d_out, 32, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    dev = d_out.cpu().numpy().view(np.uint64)
    assert np.array_equal(dev, host), "same kernels, same inputs: must be bit-identical run to run"
    assert eng.kernel_launches > 0


def test_pbs_then_keyswitch_order(orc, keys_2_2, eng):
    """PBSOrder::BootstrapKeyswitch (server_key/mod.rs:859-932): input and output under the SMALL key (LWE noise)."""
    import ctypes as C
    p, ck, sk = keys_2_2
    L = orc.lib()
    fs, luts = _luts(sk)
    eng.upload_luts(luts)
    vals = np.arange(16)
    cts = np.zeros((16, p.lwe_dim + 1), dtype=np.uint64)
    for i, v in enumerate(vals):
        L.orc_lwe_encrypt(ck.small_sk, p.lwe_dim, int(v) << 59, p.lwe_std, C.byref(ck.rng), cts[i])
    idx = (np.arange(16) % len(fs)).astype(np.uint32)
    out = eng.pbs_ks_batch(cts, idx)
    got = [L.orc_decode(C.byref(p), ck.decrypt_small_raw(o)) for o in out]
    assert got == [fs[i](int(v)) for v, i in zip(vals, idx)]
    # the same composition with the oracle: PBS then keyswitch
    ref = np.stack([sk.keyswitch(sk.pbs(c, luts[i])) for c, i in zip(cts, idx)])
    assert [L.orc_decode(C.byref(p), ck.decrypt_small_raw(o)) for o in ref] == got


def test_every_kernel_instance_boundary(orc, keys_2_2, eng):
    """Batch sizes around the switches between the 1-, 2- and 4-ciphertext-per-CTA instances of the blind-rotation kernel (and ragged
    tails inside a CTA): every ciphertext must decrypt to its LUT value, and a ciphertext's result must not depend on which instance or
    which batch it was computed in: identical words within a kernel family (narrow levels run pbs_v8.cu, wide ones pbs_v4.cu)."""
    import torch
    p, ck, sk = keys_2_2
    fs, luts = _luts(sk)
    eng.upload_luts(luts)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    base = ck.encrypt_batch(np.arange(64) % 16)                      # 64 real ciphertexts, tiled to the batch size
    first = {}
    for batch in (1, 3, sms, sms + 1, 2 * sms, 2 * sms + 1, 4 * sms + 3):
        reps = -(-batch // 64)
        cts = np.tile(base, (reps, 1))[:batch]
        vals = np.tile(np.arange(64) % 16, reps)[:batch]
        idx = (np.arange(batch) % len(fs)).astype(np.uint32)
        out = eng.ks_pbs_batch(cts, idx)
        want = np.array([fs[i](int(v)) for v, i in zip(vals, idx)])
        assert np.array_equal(ck.decrypt_batch(out), want), batch
        family = "narrow" if batch <= 2 * sms else "wide"     # pbs_v8.cu / pbs_v4.cu: different FFT factorisations, different rounding
        first.setdefault(family, out[0].copy())
        assert np.array_equal(out[0], first[family]), f"ciphertext 0 differs between batch sizes (batch {batch})"


@pytest.mark.parametrize("log2_q", [63, 32, 16])
def test_power_of_two_ciphertext_modulus(orc, keys_2_2, log2_q):
    """Non-native power-of-two ciphertext modulus (bootstrap.rs:318-330; the reference's own PBS tests run q = 2^63,
    core_crypto/algorithms/test/mod.rs TEST_PARAMS_3_BITS_63_U64): every PBS output word is a multiple of 2^(64 - log2_q); LUT rotation
    + rounding + sample extraction is bit-exact against the reference's order (round the accumulator, then extract), ties included;
    one CMUX differs by at most one unit of the modulus or the FFT tolerance; full KS-PBS decrypts exactly and equals the oracle's
    decryption."""
    import fhe_string_bounty_b200 as F
    p, ck, sk = keys_2_2
    fs, luts = _luts(sk)
    eng = F.Engine(engine_params(p))
    eng.upload_ksk(sk.ksk)
    eng.upload_bsk_std(sk.bsk)
    eng.upload_luts(luts)
    eng.set_ciphertext_modulus_log2(log2_q)
    unit = 1 << (64 - log2_q)
    shift = np.uint64(64 - log2_q - 1)
    rnd = lambda x: (((x >> shift) + np.uint64(1)) & ~np.uint64(1)) << shift
    vals = np.arange(32) % 16
    idx = (np.arange(32) % len(fs)).astype(np.uint32)
    cts = rnd(ck.encrypt_batch(vals))                    # ciphertexts modulo q live in the MSBs
    small = np.stack([sk.keyswitch(c) for c in cts])
    assert np.array_equal(eng.keyswitch_batch(cts), small)
    # an accumulator whose words sit exactly on rounding ties, so that the order "round, then negate" matters
    tie_lut = luts[0].copy()
    tie_lut[p.poly_size:] = (np.arange(p.poly_size, dtype=np.uint64) << np.uint64(64 - log2_q)) + np.uint64(unit // 2)
    eng.upload_luts(np.concatenate([luts, tie_lut[None, :]]))
    all_luts = np.concatenate([luts, tie_lut[None, :]])
    idx0 = np.array([len(fs), 0, 1, 2], dtype=np.uint32)
    for n_iters in (0, 1):
        got = eng.pbs_batch(small[:4], idx0, n_iters=n_iters)
        assert not np.any(got & np.uint64(unit - 1)), "outputs must be multiples of 2^(64 - log2_q)"
        for b in range(4):
            want = oracle_partial_pbs(orc, sk, small[b], all_luts[idx0[b]], n_iters, log2_q)
            d = int(np.abs((got[b] - want).view(np.int64)).max())
            if n_iters == 0:
                assert d == 0, (log2_q, b)
            else:
                assert d <= max(unit, 2**44), (log2_q, b, np.log2(max(d, 1)))
    out = eng.ks_pbs_batch(cts, idx)
    assert not np.any(out & np.uint64(unit - 1))
    want = np.array([fs[i](int(v)) for v, i in zip(vals, idx)])
    assert np.array_equal(ck.decrypt_batch(out), want)
    ref = np.stack([sk.pbs_pow2_modulus(small[b], luts[idx[b]], log2_q) for b in range(8)])
    assert not np.any(ref & np.uint64(unit - 1))
    assert np.array_equal(ck.decrypt_batch(ref), want[:8])
    eng.close()
