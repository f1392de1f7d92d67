"""GPU parity for the multi-bit PBS (grouping factor 3, BASELINE config 5) against the oracle's deterministic multi-bit
restatement (lwe_multi_bit_programmable_bootstrapping.rs:18-84,548-): keyswitch bit-exact (base 2^7, 2 levels), decrypted
LUT outputs bit-exact over all messages, phase error comparable to the oracle's, run-to-run determinism
(test/lwe_multi_bit_programmable_bootstrapping.rs:295), and lexicographic lt/le on strings."""
import numpy as np
import pytest

from fhe_string_bounty_b200.host import Program
from helpers import engine_params, phase_error

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng(keys_multibit):
    import fhe_string_bounty_b200 as F
    p, ck, sk = keys_multibit
    e = F.Engine(engine_params(p))
    e.upload_ksk(sk.ksk)
    e.upload_bsk_std(sk.bsk)
    yield e
    e.close()


def test_multibit_keyswitch_bit_exact(orc, keys_multibit, eng):
    p, ck, sk = keys_multibit
    rng = np.random.default_rng(50)
    cts = ck.encrypt_batch(rng.integers(0, 16, size=67))
    cts[0, :] = np.uint64(2**64 - 1)
    cts[1, :] = rng.integers(0, 2**64, size=cts.shape[1], dtype=np.uint64)
    got = eng.keyswitch_batch(cts)
    want = np.stack([sk.keyswitch(c) for c in cts])
    assert np.array_equal(got, want)


def test_multibit_pbs_all_messages(orc, keys_multibit, eng):
    p, ck, sk = keys_multibit
    fs = [lambda x: x, lambda x: (5 * x + 3) % 16, lambda x: int(x != 0)]
    luts = np.stack([sk.generate_lookup_table(f)[0] for f in fs])
    eng.upload_luts(luts)
    vals = np.array([v for v in range(16) for _ in fs])
    idx = np.array([i for _ in range(16) for i in range(len(fs))], dtype=np.uint32)
    cts = ck.encrypt_batch(vals)
    out = eng.ks_pbs_batch(cts, idx)
    want = np.array([fs[i](int(v)) for v, i in zip(vals, idx)])
    assert np.array_equal(ck.decrypt_batch(out), want)
    ref = sk.ks_pbs_batch(cts, luts, idx)
    assert np.array_equal(ck.decrypt_batch(ref), want)
    e_gpu, e_cpu = phase_error(ck, out, want), phase_error(ck, ref, want)
    print(f"multi-bit phase error: gpu max 2^{np.log2(e_gpu.max()):.1f} rms 2^{np.log2(np.sqrt((e_gpu**2).mean())):.1f}; "
          f"oracle max 2^{np.log2(e_cpu.max()):.1f} rms 2^{np.log2(np.sqrt((e_cpu**2).mean())):.1f}")
    assert e_gpu.max() < 2**55
    assert np.sqrt((e_gpu**2).mean()) < 2 * np.sqrt((e_cpu**2).mean()) + 2**40
    # deterministic: same inputs twice -> identical ciphertext words
    assert np.array_equal(out, eng.ks_pbs_batch(cts, idx))
    # zero groups processed: LUT rotation + sample extraction only, bit-exact vs the oracle's integer path
    small = np.stack([sk.keyswitch(c) for c in cts[:4]])
    got0 = eng.pbs_batch(small, idx[:4], n_iters=0)
    L = orc.lib()
    for b in range(4):
        b_hat = L.orc_modulus_switch(int(small[b][p.lwe_dim]), 11)
        acc = np.zeros(p.lut_len, dtype=np.uint64)
        for q in range(2):
            tmp = np.zeros(p.poly_size, dtype=np.uint64)
            L.orc_monomial_div(tmp, np.ascontiguousarray(luts[idx[b]][q * 2048:(q + 1) * 2048]), 2048, b_hat)
            acc[q * 2048:(q + 1) * 2048] = tmp
        want0 = np.zeros(p.big_dim + 1, dtype=np.uint64)
        import ctypes as C
        L.orc_sample_extract0(C.byref(p), acc, want0)
        assert np.array_equal(got0[b], want0)


def test_multibit_string_lt_le_config5(orc, keys_multibit, eng):
    """BASELINE config 5: lt/le on 128-char strings sharing a random-length prefix; 512 PBS in 10 levels."""
    from oracle import radix as R
    p, ck, sk = keys_multibit
    rng = np.random.default_rng(0xB200 + 5)
    progs = {op: Program("string_" + op, (128, 128), params=engine_params(p)) for op in ("lt", "le", "gt", "ge", "eq")}
    assert progs["lt"].level_widths == [256, 128, 64, 32, 16, 8, 4, 2, 1, 1]
    for trial in range(3):
        a = bytes(rng.integers(0x20, 0x7F, size=128).tolist())
        k = int(rng.integers(0, 129))
        b = a[:k] + bytes(rng.integers(0x20, 0x7F, size=128 - k).tolist()) if trial else a
        ins = np.concatenate([R.encrypt_string(ck, a), R.encrypt_string(ck, b)])
        clear = {"lt": a < b, "le": a <= b, "gt": a > b, "ge": a >= b, "eq": a == b}
        for op, P in progs.items():
            assert ck.decrypt_message_and_carry(P.run(eng, ins)[0]) == int(clear[op]), (op, k)
    print(f"lt 128 chars (multi-bit): {progs['lt'].n_pbs} PBS, {progs['lt'].last_ms():.1f} ms on device")


def test_multibit_wide_level_with_tail(orc, keys_multibit):
    """A level a little wider than one 4-ciphertext-per-SM wave: the whole wave runs on the 4-ciphertext instance, the remainder on the
    narrow instances (c_api.cu do_pbs); every ciphertext must land in its own output row and decrypt to its LUT value."""
    import torch
    import fhe_string_bounty_b200 as F
    p, ck, sk = keys_multibit
    eng = F.Engine(engine_params(p))
    eng.upload_ksk(sk.ksk)
    eng.upload_bsk_std(sk.bsk)
    acc0, _ = sk.generate_lookup_table(lambda x: (x + 1) % 16)
    acc1, _ = sk.generate_lookup_table(lambda x: (3 * x) % 16)
    eng.upload_luts(np.stack([acc0, acc1]))
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    base = ck.encrypt_batch(np.arange(48) % 16)
    for batch in (4 * sms + 7, 4 * sms + sms + 9):
        reps = -(-batch // 48)
        cts = np.tile(base, (reps, 1))[:batch]
        vals = np.tile(np.arange(48) % 16, reps)[:batch]
        idx = (np.arange(batch) % 2).astype(np.uint32)
        out = eng.ks_pbs_batch(cts, idx)
        want = [(int(v) + 1) % 16 if i == 0 else (3 * int(v)) % 16 for v, i in zip(vals, idx)]
        assert list(ck.decrypt_batch(out)) == want, batch
    eng.close()


def test_every_multibit_instance(orc, keys_multibit, monkeypatch):
    """The 1-, 2-, 3- and 4-ciphertext-per-CTA instances of the multi-bit kernel (TFHE_B200_MB_CTS forces one; by default the batch size
    picks 1, 2 or 3) on a ragged batch: same decrypted values."""
    import fhe_string_bounty_b200 as F
    p, ck, sk = keys_multibit
    acc, _ = sk.generate_lookup_table(lambda x: (9 * x + 4) % 16)
    vals = np.arange(13) % 16
    cts = ck.encrypt_batch(vals)
    want = [(9 * int(v) + 4) % 16 for v in vals]
    monkeypatch.setenv("TFHE_B200_NARROW_KERNEL", "0")           # read at context creation: keep the narrow batch on pbs_multibit_v4.cu
    eng = F.Engine(engine_params(p))
    eng.upload_ksk(sk.ksk)
    eng.upload_bsk_std(sk.bsk)
    eng.upload_luts(acc[None, :])
    outs = []
    for cts_per_cta in ("1", "2", "3", "4"):
        monkeypatch.setenv("TFHE_B200_MB_CTS", cts_per_cta)      # read by the launcher at every launch
        out = eng.ks_pbs_batch(cts, None)
        assert list(ck.decrypt_batch(out)) == want, cts_per_cta
        outs.append(out)
    eng.close()


def test_multibit_narrow_level_kernel(orc, keys_multibit, monkeypatch):
    """Levels of at most one ciphertext per SM run on pbs_multibit_v8.cu (8 FFT points per thread, accumulator in registers). Against the
    16-points-per-thread kernel on the same inputs: LUT rotation + sample extraction identical words, one multi-bit step within the FFT
    tolerance (2^44), full PBS decrypting identically over all messages, for 1 ciphertext, a ragged batch and a full
    wave; a wide level's remainder lands in the right output rows."""
    import torch
    import fhe_string_bounty_b200 as F
    p, ck, sk = keys_multibit
    fs = [lambda x: (7 * x + 2) % 16, lambda x: int(x >= 5)]
    luts = np.stack([sk.generate_lookup_table(f)[0] for f in fs])
    sms = torch.cuda.get_device_properties(0).multi_processor_count

    def make():
        e = F.Engine(engine_params(p))
        e.upload_ksk(sk.ksk)
        e.upload_bsk_std(sk.bsk)
        e.upload_luts(luts)
        return e

    e8 = make()
    monkeypatch.setenv("TFHE_B200_NARROW_KERNEL", "0")
    e4 = make()
    monkeypatch.delenv("TFHE_B200_NARROW_KERNEL")
    base = ck.encrypt_batch(np.arange(32) % 16)
    for batch in (1, 37, sms):
        reps = -(-batch // 32)
        cts = np.tile(base, (reps, 1))[:batch]
        vals = np.tile(np.arange(32) % 16, reps)[:batch]
        idx = (np.arange(batch) % 2).astype(np.uint32)
        want = [fs[i](int(v)) for v, i in zip(vals, idx)]
        out8 = e8.ks_pbs_batch(cts, idx)
        assert list(ck.decrypt_batch(out8)) == want, batch
        assert list(ck.decrypt_batch(e4.ks_pbs_batch(cts, idx))) == want, batch
        assert np.array_equal(out8, e8.ks_pbs_batch(cts, idx))          # deterministic
        if batch <= sms // 2:    # default: one ciphertext per two-SM cluster (pbs_multibit_kernel_v8x2); identical words on one SM
            e8.set_tuning("narrow_cluster", 0)
            assert np.array_equal(out8, e8.ks_pbs_batch(cts, idx)), f"cluster instance differs from the one-SM instance, batch {batch}"
            e8.set_tuning("narrow_cluster", 1)
        err = phase_error(ck, out8, np.array(want))
        assert err.max() < 2**55
        small = e8.keyswitch_batch(cts)
        assert np.array_equal(e8.pbs_batch(small, idx, n_iters=0), e4.pbs_batch(small, idx, n_iters=0))
        d = (e8.pbs_batch(small, idx, n_iters=1) - e4.pbs_batch(small, idx, n_iters=1)).view(np.int64)
        assert np.abs(d).max() <= 2**44, (batch, int(np.abs(d).max()))
    # three ciphertexts per SM + a remainder of 11: the remainder runs on the narrow kernel
    batch = 3 * sms + 11
    reps = -(-batch // 32)
    cts = np.tile(base, (reps, 1))[:batch]
    vals = np.tile(np.arange(32) % 16, reps)[:batch]
    idx = (np.arange(batch) % 2).astype(np.uint32)
    assert list(ck.decrypt_batch(e8.ks_pbs_batch(cts, idx))) == [fs[i](int(v)) for v, i in zip(vals, idx)]
    e8.close()
    e4.close()
