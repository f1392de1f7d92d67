"""tfhe-rs 0.5 wire format (bincode 1.3.3, fixed-width little-endian) for the objects that cross the KS-PBS boundary: thin ctypes wrappers
over the C ABI (include/tfhe_b200.h, csrc/host/wire.h).  Parsing needs no GPU."""
import ctypes as C

import numpy as np

from . import _native as N


class WireServerKey(C.Structure):
    _fields_ = [("params", N.Params), ("pbs_order", C.c_uint32), ("deterministic_execution", C.c_uint32), ("max_degree", C.c_uint64),
                ("ksk_seed", C.c_uint8 * 16), ("bsk_seed", C.c_uint8 * 16), ("ksk_byte_offset", C.c_uint64), ("ksk_words", C.c_uint64),
                ("bsk_byte_offset", C.c_uint64), ("bsk_words", C.c_uint64)]


def _check(lib, rc):
    if rc != 0:
        raise N.NativeError(lib.tfhe_b200_last_error().decode())


def parse_compressed_server_key(blob: bytes) -> WireServerKey:
    """-> parameter set, PBS order, seeds and the position of the two body arrays inside `blob`"""
    lib = N.load_native()
    buf = np.frombuffer(blob, dtype=np.uint8)
    out = WireServerKey()
    _check(lib, lib.tfhe_b200_wire_parse_compressed_server_key(N._ptr(buf), buf.size, C.byref(out)))
    return out


def read_ciphertexts(blob: bytes, radix: bool):
    """-> (lwe [n, lwe_len] u64, meta [n, 5] u64 = degree, noise_level, message_modulus, carry_modulus, pbs_order)"""
    lib = N.load_native()
    buf = np.frombuffer(blob, dtype=np.uint8)
    n, ll = C.c_size_t(0), C.c_size_t(0)
    _check(lib, lib.tfhe_b200_wire_read_ciphertexts(N._ptr(buf), buf.size, int(radix), None, 0, None, C.byref(n), C.byref(ll)))
    lwe = np.zeros((n.value, ll.value), dtype=np.uint64)
    meta = np.zeros((n.value, 5), dtype=np.uint64)
    _check(lib, lib.tfhe_b200_wire_read_ciphertexts(N._ptr(buf), buf.size, int(radix), N._ptr(lwe), lwe.size, N._ptr(meta), C.byref(n),
                                                    C.byref(ll)))
    return lwe, meta


def write_ciphertexts(lwe: np.ndarray, meta: np.ndarray, radix: bool) -> bytes:
    lib = N.load_native()
    lwe = np.ascontiguousarray(lwe, dtype=np.uint64)
    lwe = lwe.reshape(-1, lwe.shape[-1])
    meta = np.ascontiguousarray(meta, dtype=np.uint64).reshape(lwe.shape[0], 5)
    n = C.c_size_t(0)
    _check(lib, lib.tfhe_b200_wire_write_ciphertexts(N._ptr(lwe), lwe.shape[1], N._ptr(meta), lwe.shape[0], int(radix), None, 0, C.byref(n)))
    out = np.zeros(n.value, dtype=np.uint8)
    _check(lib, lib.tfhe_b200_wire_write_ciphertexts(N._ptr(lwe), lwe.shape[1], N._ptr(meta), lwe.shape[0], int(radix), N._ptr(out), out.size,
                                                     C.byref(n)))
    return out.tobytes()
