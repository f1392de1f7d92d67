"""Ciphertext-level parity of the kernels that carry the headline numbers (run with -m gpu).

The wide-level instances -- pbs_classic_kernel_v4<4> (four ciphertexts per SM, 95 % of the bench step), its <2> / <1> siblings and
pbs_multibit_kernel_v4<3> -- are only reached by batches wider than two ciphertexts per SM (c_api.cu do_pbs_kernels), so the small
batches of test_gpu_parity.py never touch them.  Here every instance is pinned by batch size / tfhe_b200_set_tuning and compared with the
oracle at the ciphertext level:
  * n_iters = 0: LUT rotation + sample extraction, bit-exact;
  * n_iters = 1: one CMUX (classic) / one multi-bit group step, max |delta| <= 2^44 u64 torus units vs the oracle's f64 external
    product (stated bound; the reference's own FFT tolerance at these digit sizes is 2^46, fft/tests.rs:166-167);
  * full PBS: decrypted values exact on every ciphertext of the batch, phase error max / rms next to the oracle's on the same inputs;
  * fused (u16 modulus-switched hand-off) KS -> PBS == unfused keyswitch_batch + pbs_batch, word for word.
"""
import numpy as np
import pytest

from helpers import engine_params, phase_error

pytestmark = pytest.mark.gpu

FS = [lambda x: x, lambda x: x % 4, lambda x: x // 4, lambda x: int(x == 5), lambda x: (3 * x + 1) % 16]


def _sms():
    import torch
    return torch.cuda.get_device_properties(0).multi_processor_count


def _make_engine(p, sk, luts):
    import fhe_string_bounty_b200 as F
    e = F.Engine(engine_params(p))
    e.upload_ksk(sk.ksk)
    e.upload_bsk_std(sk.bsk)
    e.upload_luts(luts)
    return e


def _batch(ck, batch, n_luts, seed):
    """`batch` ciphertexts built from 64 real encryptions (tiled), their clear values and LUT indices"""
    rng = np.random.default_rng(seed)
    base_vals = rng.integers(0, 16, size=64)
    base = ck.encrypt_batch(base_vals)
    reps = -(-batch // 64)
    cts = np.tile(base, (reps, 1))[:batch]
    vals = np.tile(base_vals, reps)[:batch]
    idx = ((np.arange(batch) * 7 + np.arange(batch) // 64) % n_luts).astype(np.uint32)
    return cts, vals, idx


def _probe_rows(batch, per_cta, n=24):
    """rows to compare with the oracle: the first and last CTA of the wide launch, every position inside a CTA, and a spread in between"""
    rows = set(range(per_cta)) | set(range(batch - per_cta, batch)) | {batch // 2 + r for r in range(per_cta)}
    rng = np.random.default_rng(batch)
    rows |= set(int(r) for r in rng.integers(0, batch, size=n))
    return sorted(r for r in rows if 0 <= r < batch)


def _check_partial(eng, sk, small, idx, luts, rows, label):
    worst = 0
    for n_iters in (0, 1):
        got = eng.pbs_batch(small, idx, n_iters=n_iters)
        for b in rows:
            want = sk.pbs_partial(small[b], luts[idx[b]], n_iters)
            d = int(np.abs((got[b] - want).view(np.int64)).max())
            if n_iters == 0:
                assert d == 0, f"{label}: LUT rotation / sample extraction must be bit-exact (row {b})"
            else:
                worst = max(worst, d)
                assert d <= 2**44, f"{label}: row {b}: one step max|delta| = 2^{np.log2(max(d, 1)):.1f}"
    print(f"{label}: max |delta| vs oracle-f64 after one step over {len(rows)} rows: 2^{np.log2(max(worst, 1)):.1f} (bound 2^44)")
    return worst


def _check_full(eng, ck, sk, cts, vals, idx, luts, fs, rows, label, bound=2**54):
    out = eng.ks_pbs_batch(cts, idx)
    want = np.array([fs[i](int(v)) for v, i in zip(vals, idx)])
    got = ck.decrypt_batch(out)
    assert np.array_equal(got, want), f"{label}: {np.argwhere(got != want)[:8].ravel()}"
    err_gpu = phase_error(ck, out, want)
    ref = sk.ks_pbs_batch(cts[rows], luts, idx[rows])
    assert np.array_equal(ck.decrypt_batch(ref), want[rows])
    err_cpu = phase_error(ck, ref, want[rows])
    rms = lambda e: float(np.sqrt((e**2).mean()))
    print(f"{label}: phase error over {len(cts)} cts: gpu max 2^{np.log2(err_gpu.max()):.1f} rms 2^{np.log2(rms(err_gpu)):.1f}; "
          f"oracle ({len(rows)} of them) max 2^{np.log2(err_cpu.max()):.1f} rms 2^{np.log2(rms(err_cpu)):.1f}")
    assert err_gpu.max() < bound                      # decoding margin 2^58
    assert rms(err_gpu) < 2 * rms(err_cpu) + 2**40
    return out


@pytest.fixture(scope="module")
def classic(keys_2_2):
    p, ck, sk = keys_2_2
    luts = np.stack([sk.generate_lookup_table(f)[0] for f in FS])
    e = _make_engine(p, sk, luts)
    yield p, ck, sk, luts, e
    e.close()


def test_classic_v4_four_per_sm(orc, classic):
    """batch = 4 * SMs + 3: pbs_classic_kernel_v4<4> (exchange A through Tensor Memory) takes the 4 * SMs wide part, the remainder of 3
    goes to pbs_v8.cu"""
    p, ck, sk, luts, eng = classic
    sms = _sms()
    per_sm = 4
    batch = per_sm * sms + 3
    cts, vals, idx = _batch(ck, batch, len(FS), 101)
    small = eng.keyswitch_batch(cts)
    assert np.array_equal(small[:64], np.stack([sk.keyswitch(c) for c in cts[:64]]))
    rows = _probe_rows(per_sm * sms, per_sm) + [batch - 3, batch - 2, batch - 1]
    _check_partial(eng, sk, small, idx, luts, rows, f"pbs_classic_kernel_v4<{per_sm}> (+ v8 tail)")
    out = _check_full(eng, ck, sk, cts, vals, idx, luts, FS, rows[:32], f"pbs_classic_kernel_v4<{per_sm}> (+ v8 tail)")
    # the fused u16 hand-off (keyswitch epilogue applies the modulus switch) against the unfused two-call path: same kernels, same words
    assert np.array_equal(out, eng.pbs_batch(small, idx)), "fused KS->PBS hand-off differs from keyswitch_batch + pbs_batch"
    # a ciphertext's result does not depend on its position in the wide launch (rows 0 and 64 hold the same input and LUT index when 64 | 7*64)
    same = [b for b in range(64, per_sm * sms) if idx[b] == idx[b % 64]][:16]
    for b in same:
        assert np.array_equal(out[b], out[b % 64]), f"row {b} differs from row {b % 64} (same input, same LUT)"


def test_classic_v4_three_per_sm(orc, classic):
    """Levels (and remainders of wider levels) of more than two but at most three ciphertexts per SM run pbs_classic_kernel_v4<3> (12 warps,
    152 registers, 7-slot ring): batch = 3 * SMs - 1 on its own, and 4 * SMs + 3 * SMs - 2 = one full wave of <4> followed by <3>."""
    p, ck, sk, luts, eng = classic
    sms = _sms()
    for batch, label in ((3 * sms - 1, "pbs_classic_kernel_v4<3>"), (7 * sms - 2, "pbs_classic_kernel_v4<4> + <3> remainder")):
        cts, vals, idx = _batch(ck, batch, len(FS), 300 + batch)
        small = eng.keyswitch_batch(cts)
        rows = _probe_rows(batch, 3, n=12) + ([4 * sms - 1, 4 * sms, 4 * sms + 1] if batch > 4 * sms else [])
        _check_partial(eng, sk, small, idx, luts, rows, label)
        out = _check_full(eng, ck, sk, cts, vals, idx, luts, FS, rows[:24], label)
        assert np.array_equal(out, eng.pbs_batch(small, idx)), "fused KS->PBS hand-off differs from keyswitch_batch + pbs_batch"
        same = [b for b in range(64, batch) if idx[b] == idx[b % 64]][:16]
        for b in same:
            assert np.array_equal(out[b], out[b % 64]), f"{label}: row {b} differs from row {b % 64} (same input, same LUT)"


def test_planned_level_cuts(orc, classic):
    """Cuts of plan_classic_level (c_api.cu) that no other test reaches: 4 SMs + 200 = two waves of the 3-per-SM instance in ONE launch
    (more CTAs than SMs), 5 SMs = a wave of three followed by a two-per-SM tail on pbs_v8.cu.  Same checks as the pinned instances, and the
    planner's cut is the one the test expects."""
    import ctypes as C
    import fhe_string_bounty_b200 as F
    p, ck, sk, luts, eng = classic
    sms = _sms()
    lib = F.load_native()
    for batch, want_cut in ((4 * sms + 200, (0, 4 * sms + 200, 0)), (5 * sms, (0, 3 * sms, 2 * sms))):
        cut = (C.c_size_t * 3)()
        lib.tfhe_b200_plan_classic_level(batch, sms, 2 * sms, 1, cut)
        assert tuple(cut) == want_cut, tuple(cut)
        label = f"level of {batch} cut {want_cut}"
        cts, vals, idx = _batch(ck, batch, len(FS), 500 + batch)
        small = eng.keyswitch_batch(cts)
        rows = _probe_rows(batch, 3, n=12) + [3 * sms - 1, 3 * sms, 3 * sms + 1]
        _check_partial(eng, sk, small, idx, luts, rows, label)
        out = _check_full(eng, ck, sk, cts, vals, idx, luts, FS, rows[:24], label)
        assert np.array_equal(out, eng.pbs_batch(small, idx)), "fused KS->PBS hand-off differs from keyswitch_batch + pbs_batch"
        wide_rows = want_cut[0] + want_cut[1]
        same = [b for b in range(64, wide_rows) if idx[b] == idx[b % 64]][:16]
        for b in same:
            assert np.array_equal(out[b], out[b % 64]), f"{label}: row {b} differs from row {b % 64} (same input, same LUT)"


@pytest.mark.parametrize("per_cta", [2, 1])
def test_classic_v4_narrow_instances(orc, classic, per_cta):
    """narrow_kernel = 0 keeps levels of <= 2 x SMs on pbs_v4.cu: <2> for SMs < batch <= 2 SMs, <1> for batch <= SMs (ragged tails)"""
    p, ck, sk, luts, eng = classic
    sms = _sms()
    batch = per_cta * sms - 1
    eng.set_tuning("narrow_kernel", 0)
    try:
        cts, vals, idx = _batch(ck, batch, len(FS), 200 + per_cta)
        small = eng.keyswitch_batch(cts)
        rows = _probe_rows(batch, per_cta, n=12)
        label = f"pbs_classic_kernel_v4<{per_cta}>"
        _check_partial(eng, sk, small, idx, luts, rows, label)
        out4 = _check_full(eng, ck, sk, cts, vals, idx, luts, FS, rows[:16], label)
        assert np.array_equal(out4, eng.pbs_batch(small, idx))
    finally:
        eng.set_tuning("narrow_kernel", 8)
    # the default narrow kernel (pbs_v8.cu) on the same batch: same decrypted values, one CMUX within the same bound of the oracle
    rows8 = rows[:8]
    _check_partial(eng, sk, small, idx, luts, rows8, f"pbs_classic_kernel_v8 (batch {batch})")
    out8 = eng.ks_pbs_batch(cts, idx)
    assert np.array_equal(ck.decrypt_batch(out8), ck.decrypt_batch(out4))


def test_set_tuning_rejects_bad_keys(classic):
    import fhe_string_bounty_b200 as F
    eng = classic[4]
    with pytest.raises(F.NativeError, match="unknown tuning key"):
        eng.set_tuning("no_such_key", 1)
    with pytest.raises(F.NativeError, match="narrow_kernel"):
        eng.set_tuning("narrow_kernel", 3)


def test_multibit_v4_three_per_sm(orc, keys_multibit):
    """batch = 3 * SMs + 2: pbs_multibit_kernel_v4<3> takes the 3 * SMs wide part, the remainder of 2 runs on pbs_multibit_v8.cu.
    One step = one group of three mask elements: G_0 + sum_j G_j X^deg_j combined in the Fourier domain, then one external product that
    REPLACES the accumulator (lwe_multi_bit_programmable_bootstrapping.rs:755-800)."""
    p, ck, sk = keys_multibit
    fs = [lambda x: x, lambda x: (5 * x + 3) % 16, lambda x: int(x != 0)]
    luts = np.stack([sk.generate_lookup_table(f)[0] for f in fs])
    eng = _make_engine(p, sk, luts)
    sms = _sms()
    batch = 3 * sms + 2
    cts, vals, idx = _batch(ck, batch, len(fs), 303)
    small = eng.keyswitch_batch(cts)
    assert np.array_equal(small[:48], np.stack([sk.keyswitch(c) for c in cts[:48]]))
    rows = _probe_rows(3 * sms, 3) + [batch - 2, batch - 1]
    _check_partial(eng, sk, small, idx, luts, rows, "pbs_multibit_kernel_v4<3> (+ v8 tail)")
    out = _check_full(eng, ck, sk, cts, vals, idx, luts, fs, rows[:24], "pbs_multibit_kernel_v4<3> (+ v8 tail)", bound=2**55)
    assert np.array_equal(out, eng.ks_pbs_batch(cts, idx)), "multi-bit PBS must be deterministic run to run"
    # 1- and 2-per-SM instances of the wide kernel
    eng.set_tuning("narrow_kernel", 0)
    for per_cta in (1, 2):
        b2 = per_cta * sms - 1
        rows2 = _probe_rows(b2, per_cta, n=8)
        _check_partial(eng, sk, small[:b2], idx[:b2], luts, rows2, f"pbs_multibit_kernel_v4<{per_cta}>")
    eng.close()
