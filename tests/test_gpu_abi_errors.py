"""Error behaviour of the C ABI on a live GPU (reference convention: non-zero return + message, c_api/utils.rs:3-28;
nothing falls back to the CPU) and edge-case batches."""
import ctypes as C

import numpy as np
import pytest

from helpers import engine_params

pytestmark = pytest.mark.gpu


def test_missing_keys_and_bad_arguments_fail_loudly(orc, keys_2_2):
    import fhe_string_bounty_b200 as F
    p, ck, sk = keys_2_2
    eng = F.Engine(engine_params(p))
    cts = ck.encrypt_batch([1, 2, 3])
    with pytest.raises(F.NativeError, match="keyswitch key not uploaded"):
        eng.keyswitch_batch(cts)
    with pytest.raises(F.NativeError, match="length"):
        eng.upload_ksk(sk.ksk[:-1])
    with pytest.raises(F.NativeError, match="length"):
        eng.upload_bsk_std(sk.bsk[: sk.bsk.size // 2])
    eng.upload_ksk(sk.ksk)
    with pytest.raises(F.NativeError, match="bootstrap key not uploaded"):
        eng.ks_pbs_batch(cts, None)
    eng.upload_bsk_std(sk.bsk)
    with pytest.raises(F.NativeError, match="lookup tables"):
        eng.ks_pbs_batch(cts, None)
    lib = eng.lib
    assert lib.tfhe_b200_ks_pbs_batch(eng.h, None, None, None, 3) != 0          # null buffers
    assert lib.tfhe_b200_ks_pbs_batch(None, cts.ctypes.data, None, cts.ctypes.data, 3) != 0   # null context
    assert lib.tfhe_b200_ks_pbs_batch(eng.h, None, None, None, 0) == 0          # empty batch is a no-op
    acc, _ = sk.generate_lookup_table(lambda x: x)
    eng.upload_luts(acc[None, :])
    with pytest.raises(F.NativeError, match="n_iters"):
        eng.pbs_batch(np.zeros((1, p.lwe_dim + 1), dtype=np.uint64), None, n_iters=p.lwe_dim + 1)
    h = C.c_void_p()
    q = F.Params(**engine_params(p))
    assert lib.tfhe_b200_ctx_create(99, C.byref(q), C.byref(h)) != 0 and h.value is None   # no such device
    eng.close()


def test_both_keyswitch_kernels_and_both_pbs_kernel_families_agree(orc, keys_2_2):
    """The IMAD keyswitch and the two tensor-core keyswitches (mma.sync, tcgen05.mma kind::i8) are bit-identical (exact integer arithmetic, different association order of
    wrapping sums); the narrow-level kernel (pbs_v8.cu, 8 x 8 x 4 x 4 FFT) and the 1-ciphertext instance of the wide kernel
    (pbs_v4.cu, 16 x 4 x 16 FFT) evaluate the same transform with different rounding: identical LUT rotation, one CMUX within the
    stated 2^44, same decrypted values.  Kernel selection through tfhe_b200_set_tuning (the wide instances are pinned in
    tests/test_gpu_wide_kernels.py)."""
    import fhe_string_bounty_b200 as F
    p, ck, sk = keys_2_2
    acc, _ = sk.generate_lookup_table(lambda x: (7 * x + 2) % 16)
    cts = ck.encrypt_batch(np.arange(37) % 16)
    eng = F.Engine(engine_params(p))
    eng.upload_ksk(sk.ksk)
    eng.upload_bsk_std(sk.bsk)
    eng.upload_luts(acc[None, :])
    outs, kss, part = {}, {}, {}
    for ks_k, narrow in ((0, 8), (1, 8), (1, 0), (2, 8)):
        eng.set_tuning("ks_kernel", ks_k)
        eng.set_tuning("narrow_kernel", narrow)
        kss[(ks_k, narrow)] = eng.keyswitch_batch(cts)
        outs[(ks_k, narrow)] = eng.ks_pbs_batch(cts, None)
        part[(ks_k, narrow)] = [eng.pbs_batch(kss[(ks_k, narrow)], None, n_iters=n) for n in (0, 1)]
    eng.close()
    assert np.array_equal(kss[(0, 8)], kss[(1, 8)]) and np.array_equal(kss[(0, 8)], np.stack([sk.keyswitch(c) for c in cts]))
    assert np.array_equal(kss[(0, 8)], kss[(2, 8)]), "tcgen05 keyswitch differs from the IMAD keyswitch"
    assert np.array_equal(outs[(1, 8)], outs[(2, 8)]), "fused u16 hand-off: tcgen05 vs mma.sync keyswitch"
    want = [(7 * int(v) + 2) % 16 for v in np.arange(37) % 16]
    for k, o in outs.items():
        assert list(ck.decrypt_batch(o)) == want, k
    # the unfused path (IMAD keyswitch, u64 hand-off) and the fused one (tensor-core keyswitch, u16 hand-off) feed the same PBS kernel
    assert np.array_equal(outs[(0, 8)], outs[(1, 8)])
    assert np.array_equal(part[(1, 8)][0], part[(1, 0)][0])
    d = np.abs((part[(1, 8)][1] - part[(1, 0)][1]).view(np.int64)).max()
    assert d <= 2**44, f"v8 vs v4<1> after one CMUX: 2^{np.log2(max(int(d), 1)):.1f}"


def test_two_sm_cluster_instance_is_bit_identical_to_the_one_sm_instance(orc, keys_2_2):
    """Levels of at most SM count / 2 ciphertexts put ONE ciphertext on a cluster of two SMs (pbs_classic_kernel_v8x2: a polynomial per
    CTA, spectra swapped through distributed shared memory).  Same FFT, same order of every floating-point operation as the one-SM
    instance pbs_classic_kernel_v8<1>: outputs are identical words, for 0 / 1 / all blind-rotation iterations, fused and unfused."""
    import fhe_string_bounty_b200 as F
    p, ck, sk = keys_2_2
    accs = np.stack([sk.generate_lookup_table(f)[0] for f in (lambda x: (3 * x + 1) % 16, lambda x: int(x >= 8))])
    eng = F.Engine(engine_params(p))
    eng.upload_ksk(sk.ksk)
    eng.upload_bsk_std(sk.bsk)
    eng.upload_luts(accs)
    for batch in (1, 37, 74):
        cts = ck.encrypt_batch(np.arange(batch) % 16)
        idx = (np.arange(batch) % 2).astype(np.uint32)
        res = {}
        for cl in (0, 1):
            eng.set_tuning("narrow_cluster", cl)
            small = eng.keyswitch_batch(cts)
            res[cl] = [eng.pbs_batch(small, idx, n_iters=n) for n in (0, 1, None)] + [eng.ks_pbs_batch(cts, idx)]
        for a, b in zip(res[0], res[1]):
            assert np.array_equal(a, b), f"batch {batch}"
        want = [[(3 * int(v) + 1) % 16, int(v >= 8)][i] for v, i in zip(np.arange(batch) % 16, idx)]
        assert list(ck.decrypt_batch(res[1][3])) == want
    eng.close()


def test_kernels_do_not_write_outside_their_buffers(orc, keys_2_2):
    """compute-sanitizer is closed on this pool, so bounds are checked with canaries: device input/output/small buffers are
    embedded in larger sentinel-filled allocations; after KS, PBS and KS-PBS on ragged batch sizes the sentinels must be intact
    and the inputs unchanged."""
    import torch
    import fhe_string_bounty_b200 as F
    p, ck, sk = keys_2_2
    eng = F.Engine(engine_params(p))
    eng.upload_ksk(sk.ksk)
    eng.upload_bsk_std(sk.bsk)
    acc, _ = sk.generate_lookup_table(lambda x: x)
    eng.upload_luts(np.stack([acc, acc]))
    SENT = -0x0123456789ABCDEF
    guard = 4096
    s = torch.cuda.Stream()
    for batch in (1, 5, 149, 601):
        cts = ck.encrypt_batch(np.arange(batch) % 16)
        n_in, n_out, n_small = batch * p.big_dim + batch, batch * (p.big_dim + 1), batch * (p.lwe_dim + 1)
        buf_in = torch.full((guard + n_in + guard,), SENT, dtype=torch.int64, device="cuda")
        buf_out = torch.full((guard + n_out + guard,), SENT, dtype=torch.int64, device="cuda")
        buf_small = torch.full((guard + n_small + guard,), SENT, dtype=torch.int64, device="cuda")
        idx = torch.zeros(guard + batch + guard, dtype=torch.int32, device="cuda")
        buf_in[guard:guard + n_in] = torch.from_numpy(cts.view(np.int64).ravel()).cuda()
        d_in, d_out, d_small = buf_in[guard:], buf_out[guard:], buf_small[guard:]
        eng.ks_pbs_batch_device(d_in, idx[guard:], d_out, batch, s.cuda_stream)
        eng.keyswitch_batch_device(d_in, d_small, batch, s.cuda_stream)
        s.synchronize()
        for name, b, n in (("in", buf_in, n_in), ("out", buf_out, n_out), ("small", buf_small, n_small)):
            assert bool((b[:guard] == SENT).all()) and bool((b[guard + n:] == SENT).all()), (name, batch)
        assert np.array_equal(buf_in[guard:guard + n_in].cpu().numpy().view(np.uint64).reshape(batch, -1), cts), "inputs are read-only"
        got = buf_out[guard:guard + n_out].cpu().numpy().view(np.uint64).reshape(batch, -1)
        assert list(ck.decrypt_batch(got)) == list(np.arange(batch) % 16)
        ks = buf_small[guard:guard + n_small].cpu().numpy().view(np.uint64).reshape(batch, -1)
        assert np.array_equal(ks[0], sk.keyswitch(cts[0]))
        eng.pbs_batch_device(d_small, idx[guard:], d_out, batch, s.cuda_stream)
        s.synchronize()
        assert bool((buf_out[:guard] == SENT).all()) and bool((buf_out[guard + n_out:] == SENT).all())
    eng.close()


def test_narrow_level_kernel_on_and_off(orc, keys_2_2, monkeypatch):
    """pbs_v8.cu (8 FFT points per thread, 8 warps per ciphertext; TFHE_B200_NARROW_KERNEL=0 falls back to pbs_v4's narrow instances) decrypts
    correctly for both of its instances (<= SM count: 1 ciphertext per CTA, registers only; <= 2 * SM count: 2 per CTA, TMEM)."""
    import torch
    import fhe_string_bounty_b200 as F
    p, ck, sk = keys_2_2
    acc, _ = sk.generate_lookup_table(lambda x: (11 * x + 5) % 16)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    base = ck.encrypt_batch(np.arange(32) % 16)
    for mode in ("8", "0"):
        monkeypatch.setenv("TFHE_B200_NARROW_KERNEL", mode)
        eng = F.Engine(engine_params(p))
        eng.upload_ksk(sk.ksk)
        eng.upload_bsk_std(sk.bsk)
        eng.upload_luts(acc[None, :])
        for batch in (1, 7, sms, sms + 5, 2 * sms):
            reps = -(-batch // 32)
            cts = np.tile(base, (reps, 1))[:batch]
            vals = np.tile(np.arange(32) % 16, reps)[:batch]
            out = eng.ks_pbs_batch(cts, None)
            assert list(ck.decrypt_batch(out)) == [(11 * int(v) + 5) % 16 for v in vals], (mode, batch)
        eng.close()
