#!/usr/bin/env python
"""Regenerates tests/golden/wire_golden.json: a serialized toy-parameter shortint::CompressedServerKey (digest + the scalar fields a parser
must recover) and a serialized 2-block radix ciphertext (in full), both written by the ORACLE's writer (oracle/wire.py), for the product's
parser (csrc/host/wire.h) to read.  No tfhe-rs build exists in this image, so the layout is restated from the reference's serde derives
("parity unpinned"); what these fixtures pin is that the product parser and the oracle writer agree, round after round."""
import hashlib
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from oracle import oracle as O  # noqa: E402
from oracle import wire as W  # noqa: E402


def build():
    p = O.params("toy")
    ck = O.ClientKey(p, 0xB200)
    sk = O.ServerKey(ck, 0xB201, fourier=False)
    csk = O.CompressedServerKey(ck, sk, ksk_seed=0x000102030405060708090A0B0C0D0E0F, bsk_seed=0xB200)
    blob = W.serialize_compressed_server_key(csk)
    cts = ck.encrypt_batch([2, 1])
    radix = W.serialize_radix(cts, degree=p.msg_mod - 1, noise_level=1, msg_mod=p.msg_mod, carry_mod=p.carry_mod)
    return {
        "server_key_sha256": hashlib.sha256(blob).hexdigest(), "server_key_len": len(blob),
        "server_key_head_hex": blob[:8].hex(), "server_key_tail_hex": blob[-60:].hex(),
        "params": {k: int(getattr(p, k)) for k in ("lwe_dim", "glwe_dim", "poly_size", "pbs_base_log", "pbs_level", "ks_base_log", "ks_level",
                                                   "grouping_factor", "msg_mod", "carry_mod")},
        "ksk_seed_hex": csk.ksk_seed.tobytes().hex(), "bsk_seed_hex": csk.bsk_seed.tobytes().hex(),
        "radix_hex": radix.hex(), "radix_decrypts_to": [2, 1],
    }


if __name__ == "__main__":
    out = Path(__file__).with_name("wire_golden.json")
    out.write_text(json.dumps(build(), indent=1) + "\n")
    print("wrote", out)
