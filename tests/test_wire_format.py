"""tfhe-rs wire format: the product's bincode reader/writer (csrc/host/wire.h, through the C ABI) against the
oracle's independent writer/reader (oracle/wire.py) and the committed fixture.  Parsing needs no GPU."""
import ctypes as C
import hashlib
import json
import struct
from pathlib import Path

import numpy as np
import pytest

GOLD = json.loads((Path(__file__).parent / "golden" / "wire_golden.json").read_text())


@pytest.fixture(scope="module")
def toy_blob(orc, toy_keys):
    from oracle import wire as W
    p, ck, sk = toy_keys
    csk = orc.CompressedServerKey(ck, sk, ksk_seed=0x000102030405060708090A0B0C0D0E0F, bsk_seed=0xB200)
    return p, ck, csk, W.serialize_compressed_server_key(csk)


def test_server_key_blob_matches_fixture_and_parses(toy_blob):
    from fhe_string_bounty_b200 import wire
    p, ck, csk, blob = toy_blob
    assert len(blob) == GOLD["server_key_len"] and hashlib.sha256(blob).hexdigest() == GOLD["server_key_sha256"]
    assert blob[:8].hex() == GOLD["server_key_head_hex"] and blob[-60:].hex() == GOLD["server_key_tail_hex"]
    v = wire.parse_compressed_server_key(blob)
    for k, want in GOLD["params"].items():
        assert getattr(v.params, k) == want, k
    assert bytes(v.ksk_seed).hex() == GOLD["ksk_seed_hex"] and bytes(v.bsk_seed).hex() == GOLD["bsk_seed_hex"]
    assert v.pbs_order == 0 and v.max_degree == p.msg_mod * p.carry_mod - 1
    # the body arrays are where the view says they are
    ksk = np.frombuffer(blob, dtype="<u8", count=v.ksk_words, offset=v.ksk_byte_offset)
    bsk = np.frombuffer(blob, dtype="<u8", count=v.bsk_words, offset=v.bsk_byte_offset)
    assert np.array_equal(ksk, csk.ksk_bodies) and np.array_equal(bsk, csk.bsk_bodies)


def test_multibit_server_key_parses(orc, keys_multibit):
    from fhe_string_bounty_b200 import wire
    from oracle import wire as W
    p, ck, sk = keys_multibit

    class Fake:      # bodies of the right size are enough for the parser; no need to compress 155 MB here
        pass
    f = Fake()
    f.p, f.ksk_seed, f.bsk_seed = p, orc.seed_bytes(1), orc.seed_bytes(2)
    L = orc.lib()
    f.ksk_bodies = np.arange(L.orc_seeded_ksk_len(C.byref(p)), dtype=np.uint64)
    f.bsk_bodies = np.zeros(L.orc_seeded_bsk_len(C.byref(p)), dtype=np.uint64)
    v = wire.parse_compressed_server_key(W.serialize_compressed_server_key(f, deterministic_execution=True))
    assert v.params.grouping_factor == 3 and v.params.lwe_dim == 888 and v.deterministic_execution == 1
    assert v.ksk_words == f.ksk_bodies.size and v.bsk_words == f.bsk_bodies.size


def test_malformed_server_key_blobs_are_rejected(toy_blob):
    from fhe_string_bounty_b200 import wire, NativeError
    _, _, _, blob = toy_blob
    with pytest.raises(NativeError, match="truncated"):
        wire.parse_compressed_server_key(blob[:-3])
    with pytest.raises(NativeError, match="trailing"):
        wire.parse_compressed_server_key(blob + b"\x00")
    with pytest.raises(NativeError, match="sequence length"):
        wire.parse_compressed_server_key(struct.pack("<Q", 1 << 60) + blob[8:])
    bad_variant = bytearray(blob)
    v = wire.parse_compressed_server_key(blob)
    variant_at = v.ksk_byte_offset + 8 * v.ksk_words + 24 + 16 + 24      # after data, 3 usizes, seed, modulus
    assert struct.unpack_from("<I", blob, variant_at)[0] == 0
    struct.pack_into("<I", bad_variant, variant_at, 7)
    with pytest.raises(NativeError, match="variant"):
        wire.parse_compressed_server_key(bytes(bad_variant))
    bad_bits = bytearray(blob)
    struct.pack_into("<Q", bad_bits, variant_at - 8, 32)                  # scalar_bits of the keyswitch key's modulus
    with pytest.raises(NativeError, match="64-bit"):
        wire.parse_compressed_server_key(bytes(bad_bits))
    bad_mod = bytearray(blob)
    bad_mod[variant_at - 24] = 1                                          # a custom (non-native) modulus
    with pytest.raises(NativeError, match="native"):
        wire.parse_compressed_server_key(bytes(bad_mod))


def test_radix_ciphertext_round_trip(orc, toy_keys):
    from fhe_string_bounty_b200 import wire
    from oracle import wire as W
    p, ck, sk = toy_keys
    blob = bytes.fromhex(GOLD["radix_hex"])
    lwe, meta = wire.read_ciphertexts(blob, radix=True)
    assert lwe.shape == (2, p.big_dim + 1)
    assert [ck.decrypt(c) for c in lwe] == GOLD["radix_decrypts_to"]
    assert meta.tolist() == [[p.msg_mod - 1, 1, p.msg_mod, p.carry_mod, 0]] * 2
    # product writer == oracle writer, and the oracle's independent reader accepts it
    again = wire.write_ciphertexts(lwe, meta, radix=True)
    assert again == blob
    lwe2, meta2 = W.deserialize_radix(again)
    assert np.array_equal(lwe2, lwe) and np.array_equal(meta2, meta)
    # a single shortint::Ciphertext
    one = W.serialize_ciphertext(lwe[0], 3, 1, p.msg_mod, p.carry_mod)
    l1, m1 = wire.read_ciphertexts(one, radix=False)
    assert np.array_equal(l1[0], lwe[0]) and wire.write_ciphertexts(l1, m1, radix=False) == one
