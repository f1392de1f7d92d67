"""TEST INFRASTRUCTURE ONLY (see oracle/tfhe_oracle.h).  Writer for the tfhe-rs 0.5 wire format, restated independently of the product's
parser from the serde derives of the reference, so that tests have blobs to feed it:

  bincode 1.3.3 `serialize` = DefaultOptions + fixint: integers little endian at their declared width, usize as u64, sequences and strings
  with a u64 length prefix, enum variants as a u32 index, bool as one byte, structs = their fields in declaration order
  (tfhe/src/safe_deserialization.rs:13-34 builds the same options by hand).

  shortint::CompressedServerKey                shortint/server_key/compressed.rs:10-17,43-55
  SeededLweKeyswitchKey                        core_crypto/entities/seeded_lwe_keyswitch_key.rs:10-21
  SeededLweBootstrapKey / ...MultiBit...       entities/seeded_lwe_bootstrap_key.rs:15-24, seeded_lwe_multi_bit_bootstrap_key.rs:15-25
  SeededGgswCiphertextList                     entities/seeded_ggsw_ciphertext_list.rs:11-23
  CompressionSeed / Seed(u128)                 commons/math/random/generator.rs:27-39
  CiphertextModulus                            commons/ciphertext_modulus.rs:41-64   (modulus: u128 with 0 = native, scalar_bits: usize)
  shortint::Ciphertext                         shortint/ciphertext/mod.rs:261-270;  LweCiphertext entities/lwe_ciphertext.rs:500-507
  BaseRadixCiphertext { blocks }               integer/ciphertext/mod.rs:18-30
  PBSOrder                                     core_crypto/commons/parameters.rs:233-245

No tfhe-rs build exists here, so these layouts are "parity unpinned": derived from the source, never checked against a real blob."""
import struct

import numpy as np


def _u32(v): return struct.pack("<I", int(v))
def _u64(v): return struct.pack("<Q", int(v))
def _u128_bytes(b16): return bytes(np.asarray(b16, dtype=np.uint8).tobytes())
def _vec_u64(a): a = np.ascontiguousarray(a, dtype="<u8"); return _u64(a.size) + a.tobytes()
def _native_modulus(): return b"\x00" * 16 + _u64(64)


def _seeded_ggsw_list(p, bodies, seed16):
    return (_vec_u64(bodies) + _u64(p.glwe_dim + 1) + _u64(p.poly_size) + _u64(p.pbs_base_log) + _u64(p.pbs_level)
            + _u128_bytes(seed16) + _native_modulus())


def serialize_compressed_server_key(csk, max_degree=None, pbs_order=0, deterministic_execution=True) -> bytes:
    """csk: oracle.CompressedServerKey"""
    p = csk.p
    out = (_vec_u64(csk.ksk_bodies) + _u64(p.ks_base_log) + _u64(p.ks_level) + _u64(p.lwe_dim + 1) + _u128_bytes(csk.ksk_seed)
           + _native_modulus())
    if p.grouping_factor == 0:
        out += _u32(0) + _seeded_ggsw_list(p, csk.bsk_bodies, csk.bsk_seed)
    else:
        out += _u32(1) + _seeded_ggsw_list(p, csk.bsk_bodies, csk.bsk_seed) + _u64(p.grouping_factor) + bytes([1 if deterministic_execution else 0])
    if max_degree is None:
        max_degree = p.msg_mod * p.carry_mod - 1          # shortint/engine/server_side.rs: MaxDegree::from_msg_carry_modulus
    out += _u64(p.msg_mod) + _u64(p.carry_mod) + _u64(max_degree) + _native_modulus() + _u32(pbs_order)
    return out


def serialize_ciphertext(ct, degree, noise_level, msg_mod, carry_mod, pbs_order=0) -> bytes:
    return _vec_u64(ct) + _native_modulus() + _u64(degree) + _u64(noise_level) + _u64(msg_mod) + _u64(carry_mod) + _u32(pbs_order)


def serialize_radix(cts, degree, noise_level, msg_mod, carry_mod, pbs_order=0) -> bytes:
    cts = np.asarray(cts, dtype=np.uint64)
    return _u64(cts.shape[0]) + b"".join(serialize_ciphertext(c, degree, noise_level, msg_mod, carry_mod, pbs_order) for c in cts)


def deserialize_radix(blob: bytes):
    """independent reader (checks the product's writer): -> (lwe [n, lwe_len], meta [n, 5])"""
    off = 0
    (n,) = struct.unpack_from("<Q", blob, off); off += 8
    lwe, meta = [], []
    for _ in range(n):
        (ll,) = struct.unpack_from("<Q", blob, off); off += 8
        lwe.append(np.frombuffer(blob, dtype="<u8", count=ll, offset=off).copy()); off += 8 * ll
        assert blob[off:off + 16] == b"\x00" * 16 and struct.unpack_from("<Q", blob, off + 16)[0] == 64
        off += 24
        d, nl, mm, cm, po = struct.unpack_from("<QQQQI", blob, off); off += 36
        meta.append((d, nl, mm, cm, po))
    assert off == len(blob)
    return np.stack(lwe), np.array(meta, dtype=np.uint64)
