"""Times the key upload paths: standard keys (host arrays -> device, repack / Fourier conversion) vs seeded keys (bodies only + device-side
AES-128 CTR mask regeneration).  Random words: the work does not depend on the key values."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import fhe_string_bounty_b200 as F

p = F.Params(**F.PARAM_MESSAGE_2_CARRY_2_KS_PBS)
eng = F.Engine(p)
rng = np.random.default_rng(1)
ksk = rng.integers(0, 2**64, size=p.ksk_len, dtype=np.uint64)
bsk = rng.integers(0, 2**64, size=p.bsk_len, dtype=np.uint64)
ksk_bodies = rng.integers(0, 2**64, size=p.poly_size * p.glwe_dim * p.ks_level, dtype=np.uint64)
bsk_bodies = rng.integers(0, 2**64, size=p.lwe_dim * p.pbs_level * (p.glwe_dim + 1) * p.poly_size, dtype=np.uint64)
seed = np.arange(16, dtype=np.uint8)


def t(f, *a):
    f(*a)
    t0 = time.perf_counter()
    for _ in range(3):
        f(*a)
    return (time.perf_counter() - t0) / 3 * 1e3


print(f"upload_ksk         {ksk.nbytes / 1e6:6.1f} MB from host  {t(eng.upload_ksk, ksk):7.2f} ms")
print(f"upload_seeded_ksk  {ksk_bodies.nbytes / 1e6:6.1f} MB from host  {t(eng.upload_seeded_ksk, seed, ksk_bodies):7.2f} ms")
print(f"upload_bsk_std     {bsk.nbytes / 1e6:6.1f} MB from host  {t(eng.upload_bsk_std, bsk):7.2f} ms")
print(f"upload_seeded_bsk  {bsk_bodies.nbytes / 1e6:6.1f} MB from host  {t(eng.upload_seeded_bsk, seed, bsk_bodies):7.2f} ms")
