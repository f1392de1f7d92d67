"""GPU parity for seeded (compressed) server keys: masks re-drawn on the device from the compression seed must reproduce, bit for bit, the
keys that the reference's decompression (oracle/csprng_oracle.c) produces."""
import ctypes as C

import numpy as np
import pytest

from helpers import engine_params

pytestmark = pytest.mark.gpu


def _engines(F, p, sk_like, csk):
    ksk, bsk = csk.decompress()                      # oracle = reference algorithm on the CPU
    a = F.Engine(engine_params(p))                   # device-side decompression
    a.upload_seeded_ksk(csk.ksk_seed, csk.ksk_bodies)
    a.upload_seeded_bsk(csk.bsk_seed, csk.bsk_bodies)
    b = F.Engine(engine_params(p))                   # standard upload of the oracle-decompressed keys
    b.upload_ksk(ksk)
    b.upload_bsk_std(bsk)
    return a, b, ksk, bsk


def test_seeded_keys_2_2_bit_exact(orc, keys_2_2):
    import fhe_string_bounty_b200 as F
    p, ck, sk = keys_2_2
    csk = orc.CompressedServerKey(ck, sk, ksk_seed=0x00112233445566778899AABBCCDDEEFF, bsk_seed=0xB200B200B200B200B200B200B200B200)
    a, b, ksk, bsk = _engines(F, p, sk, csk)
    acc, _ = sk.generate_lookup_table(lambda x: (5 * x + 3) % 16)
    for e in (a, b):
        e.upload_luts(acc[None, :])
    vals = np.arange(48) % 16
    cts = ck.encrypt_batch(vals)
    ks_a, ks_b = a.keyswitch_batch(cts), b.keyswitch_batch(cts)
    # keyswitch is exact integer arithmetic over EVERY word of the key: equal outputs <=> equal keys (up to 2^-64 collisions)
    assert np.array_equal(ks_a, ks_b)
    sk2 = orc.ServerKey.__new__(orc.ServerKey)
    sk2.p, sk2.ksk, sk2.bsk, sk2._fourier = p, ksk, bsk, None
    assert np.array_equal(ks_a[:4], np.stack([sk2.keyswitch(c) for c in cts[:4]]))
    out_a, out_b = a.ks_pbs_batch(cts, None), b.ks_pbs_batch(cts, None)
    assert np.array_equal(out_a, out_b)              # same kernel, same key bits => identical ciphertexts
    assert list(ck.decrypt_batch(out_a)) == [(5 * int(v) + 3) % 16 for v in vals]
    a.close(); b.close()


def test_seeded_keys_multibit_bit_exact(orc, keys_multibit):
    import fhe_string_bounty_b200 as F
    p, ck, sk = keys_multibit
    csk = orc.CompressedServerKey(ck, sk, ksk_seed=7, bsk_seed=(1 << 127) | 12345)
    a, b, _, _ = _engines(F, p, sk, csk)
    acc, _ = sk.generate_lookup_table(lambda x: (x * x) % 16)
    for e in (a, b):
        e.upload_luts(acc[None, :])
    vals = np.arange(20) % 16
    cts = ck.encrypt_batch(vals)
    assert np.array_equal(a.keyswitch_batch(cts), b.keyswitch_batch(cts))
    out_a, out_b = a.ks_pbs_batch(cts, None), b.ks_pbs_batch(cts, None)
    assert np.array_equal(out_a, out_b)
    assert list(ck.decrypt_batch(out_a)) == [(int(v) * int(v)) % 16 for v in vals]
    a.close(); b.close()


def test_seeded_upload_errors(orc, keys_2_2):
    import fhe_string_bounty_b200 as F
    p, ck, sk = keys_2_2
    e = F.Engine(engine_params(p))
    seed = orc.seed_bytes(1)
    with pytest.raises(RuntimeError, match="seeded keyswitch key length"):
        e.upload_seeded_ksk(seed, np.zeros(17, dtype=np.uint64))
    with pytest.raises(RuntimeError, match="seeded bootstrap key length"):
        e.upload_seeded_bsk(seed, np.zeros(2048, dtype=np.uint64))
    assert e.lib.tfhe_b200_upload_seeded_ksk(e.h, None, None, 0) != 0
    e.close()


def test_load_serialized_compressed_server_key(orc, keys_2_2):
    """bincode blob -> tfhe_b200_load_compressed_server_key == upload_seeded_* with the same seeds and bodies (identical ciphertexts);
    a blob of another parameter set is refused."""
    import fhe_string_bounty_b200 as F
    from fhe_string_bounty_b200 import wire
    from oracle import wire as W
    p, ck, sk = keys_2_2
    csk = orc.CompressedServerKey(ck, sk, ksk_seed=11, bsk_seed=22)
    blob = W.serialize_compressed_server_key(csk)
    v = wire.parse_compressed_server_key(blob)
    a = F.Engine(F.Params(**{k: getattr(v.params, k) for k, _ in F.Params._fields_}))   # parameter set discovered from the blob
    a.load_compressed_server_key(blob)
    b = F.Engine(engine_params(p))
    b.upload_seeded_ksk(csk.ksk_seed, csk.ksk_bodies)
    b.upload_seeded_bsk(csk.bsk_seed, csk.bsk_bodies)
    acc, _ = sk.generate_lookup_table(lambda x: (x + 7) % 16)
    for e in (a, b):
        e.upload_luts(acc[None, :])
    vals = np.arange(16)
    cts = ck.encrypt_batch(vals)
    out_a, out_b = a.ks_pbs_batch(cts, None), b.ks_pbs_batch(cts, None)
    assert np.array_equal(out_a, out_b)
    assert list(ck.decrypt_batch(out_a)) == [(int(x) + 7) % 16 for x in vals]
    toy = orc.params("toy")
    tck = orc.ClientKey(toy, 1)
    tsk = orc.ServerKey(tck, 2, fourier=False)
    with pytest.raises(RuntimeError, match="parameter set"):
        a.load_compressed_server_key(W.serialize_compressed_server_key(orc.CompressedServerKey(tck, tsk, 1, 2)))
    a.close(); b.close()
