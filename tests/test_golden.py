"""Golden fixtures (tests/golden/kspbs_golden.json, made by tests/golden/make_golden.py from fixed seeds): the oracle must
reproduce them bit for bit (CPU), and the CUDA keyswitch must hit the same digest on the same seeded inputs (GPU)."""
import json
import sys
from pathlib import Path

import numpy as np
import pytest

GOLD = json.loads((Path(__file__).resolve().parent / "golden" / "kspbs_golden.json").read_text())
sys.path.insert(0, str(Path(__file__).resolve().parent / "golden"))


def test_oracle_reproduces_golden(orc):
    import make_golden
    now = make_golden.build()
    assert now == GOLD


@pytest.mark.gpu
@pytest.mark.parametrize("name,fix", [("2_2", "keys_2_2"), ("multibit_2_2_g3", "keys_multibit")])
def test_gpu_keyswitch_hits_golden_digest(orc, request, name, fix):
    import fhe_string_bounty_b200 as F
    import make_golden
    from helpers import engine_params
    p, ck, sk = request.getfixturevalue(fix)
    # same seeds as the fixture => same keys; re-derive the golden inputs with a fresh client key (the session key's RNG has advanced)
    ck2 = orc.ClientKey(p, {"2_2": 0xB200 + 1, "multibit_2_2_g3": 0xB200 + 5}[name])
    cts = ck2.encrypt_batch([3, 0, 15, 8, 5])
    assert make_golden.digest(cts) == GOLD[name]["cts_sha256"]
    eng = F.Engine(engine_params(p))
    eng.upload_ksk(sk.ksk)
    ks = eng.keyswitch_batch(cts)
    eng.close()
    assert make_golden.digest(ks) == GOLD[name]["keyswitch_sha256"]
    assert [int(v) for v in ks[:, :3].ravel()] == GOLD[name]["keyswitch_first_words"]
