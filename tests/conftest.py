import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def toy_keys(orc):
    p = orc.params("toy")
    ck = orc.ClientKey(p, 0xB200)
    sk = orc.ServerKey(ck, 0xB201)
    return p, ck, sk


@pytest.fixture(scope="session")
def keys_2_2(orc):
    """PARAM_MESSAGE_2_CARRY_2_KS_PBS keys, seeded (seed 0xB200 + config index)."""
    p = orc.params("2_2")
    ck = orc.ClientKey(p, 0xB200 + 1)
    sk = orc.ServerKey(ck, 0xB300 + 1)
    return p, ck, sk


@pytest.fixture(scope="session")
def keys_multibit(orc):
    """PARAM_MULTI_BIT_MESSAGE_2_CARRY_2_GROUP_3_KS_PBS keys (shortint/parameters/multi_bit.rs:173-190), seeded."""
    p = orc.params("multibit_2_2_g3")
    ck = orc.ClientKey(p, 0xB200 + 5)
    sk = orc.ServerKey(ck, 0xB300 + 5)
    return p, ck, sk
