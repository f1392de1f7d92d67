#!/usr/bin/env python
"""Regenerates tests/golden/kspbs_golden.json.

The reference (Rust, tfhe-rs 0.5.0) cannot be built or imported in this image and holds no ciphertext-level golden
vectors for the KS-PBS path, so these fixtures are produced by the CPU ORACLE from fixed seeds: they pin
the integer arithmetic (keyswitch outputs, LUT accumulators, decomposition digits, leveled ops) bit for bit across
rounds, for the oracle AND -- through the same seeds -- for the CUDA path (tests/test_golden.py).  Only digests of the
large arrays are stored; the small ones are stored in full."""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from oracle import oracle as O  # noqa: E402


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def build():
    out = {}
    for name, seed_ck, seed_sk in (("toy", 0xB200, 0xB201), ("2_2", 0xB200 + 1, 0xB300 + 1), ("multibit_2_2_g3", 0xB200 + 5, 0xB300 + 5)):
        p = O.params(name)
        ck = O.ClientKey(p, seed_ck)
        sk = O.ServerKey(ck, seed_sk, fourier=False)
        cts = ck.encrypt_batch([3, 0, 15, 8, 5])
        ks = np.stack([sk.keyswitch(c) for c in cts])
        acc, deg = sk.generate_lookup_table(lambda x: (3 * x + 1) % 16)
        biv, _ = sk.generate_lookup_table_bivariate(lambda x, y: (2 * x * y) % 4)
        out[name] = {
            "ksk_sha256": digest(sk.ksk), "bsk_sha256": digest(sk.bsk), "cts_sha256": digest(cts),
            "keyswitch_sha256": digest(ks), "keyswitch_first_words": [int(v) for v in ks[:, :3].ravel()],
            "lut_sha256": digest(acc), "lut_degree": int(deg), "bivariate_lut_sha256": digest(biv),
            "decrypted": [int(v) for v in ck.decrypt_batch(cts)],
        }
    out["decompose"] = {f"{x}:{bl}:{lv}": O.decompose(x, bl, lv) for x, bl, lv in
                        [(0x123456789ABCDEF0, 3, 5), (2**64 - 1, 3, 5), (2**63, 23, 1), (0xDEADBEEFCAFEF00D, 7, 2), (1 << 40, 23, 1), (12345, 21, 1)]}
    return out


if __name__ == "__main__":
    path = Path(__file__).resolve().parent / "kspbs_golden.json"
    path.write_text(json.dumps(build(), indent=1, sort_keys=True) + "\n")
    print("wrote", path)
