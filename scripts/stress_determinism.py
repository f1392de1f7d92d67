"""Determinism stress: the kernels have no data-dependent scheduling, so repeated runs of one program on the same inputs must give
identical words.  Any difference is a race (ring / hand-off / exchange protocol).  usage: stress_determinism.py [reps] [multibit]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import fhe_string_bounty_b200 as F
from fhe_string_bounty_b200.host import Program

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
mb = len(sys.argv) > 2 and sys.argv[2] == "multibit"
p = F.Params(**(F.PARAM_MULTI_BIT_MESSAGE_2_CARRY_2_GROUP_3_KS_PBS if mb else F.PARAM_MESSAGE_2_CARRY_2_KS_PBS))
eng = F.Engine(p)
rng = np.random.default_rng(1)
eng.upload_ksk(rng.integers(0, 2**64, size=p.ksk_len, dtype=np.uint64))
eng.upload_bsk_std(rng.integers(0, 2**64, size=p.bsk_len, dtype=np.uint64))
bad_total = 0
for op, args in [("shortint_apply_lut", [1] + list(range(16))), ("shortint_apply_lut", [7] + list(range(16))), ("shortint_apply_lut", [74] + list(range(16))),
                 ("shortint_apply_lut", [100] + list(range(16))), ("shortint_apply_lut", [200] + list(range(16))), ("string_eq", (8, 8)),
                 ("string_find", (40, 6)), ("string_lt", (32, 32))]:
    P = Program(op, args, params=p)
    ins = rng.integers(0, 2**64, size=(P.n_inputs, p.big_len), dtype=np.uint64)
    ref = P.run(eng, ins).copy()
    bad = 0
    for r in range(reps):
        out = P.run(eng, ins)
        if not np.array_equal(out, ref):
            bad += 1
    bad_total += bad
    print(f"{op:22s} {str(tuple(args)[:2]):10s} levels {P.level_widths}: {bad} of {reps} runs differ", flush=True)
print("TOTAL differing runs:", bad_total)
