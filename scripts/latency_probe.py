"""Latency probe for small programs (device ms vs wall ms)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import fhe_string_bounty_b200 as F
from fhe_string_bounty_b200.host import Program

p = F.Params(**F.PARAM_MESSAGE_2_CARRY_2_KS_PBS)
eng = F.Engine(p)
rng = np.random.default_rng(1)
eng.upload_ksk(rng.integers(0, 2**64, size=p.ksk_len, dtype=np.uint64))
eng.upload_bsk_std(rng.integers(0, 2**64, size=p.bsk_len, dtype=np.uint64))
eng.upload_luts(rng.integers(0, 2**64, size=(4, p.lut_len), dtype=np.uint64))
import os
if os.environ.get("NARROW_CLUSTER"): eng.set_tuning("narrow_cluster", int(os.environ["NARROW_CLUSTER"]))
for op, args in [("shortint_apply_lut", [1] + list(range(16))), ("shortint_apply_lut", [32] + list(range(16))), ("shortint_apply_lut", [74] + list(range(16))),
                 ("shortint_apply_lut", [148] + list(range(16))), ("shortint_apply_lut", [592] + list(range(16))),
                 ("string_eq", (8, 8)), ("string_lt", (128, 128)), ("string_contains", (256, 16))]:
    P = Program(op, args)
    ins = rng.integers(0, 2**64, size=(P.n_inputs, p.big_len), dtype=np.uint64)
    P.run(eng, ins)
    t0 = time.perf_counter()
    for _ in range(5):
        P.run(eng, ins)
    wall = (time.perf_counter() - t0) / 5 * 1e3
    print(f"{op:22s} {str(tuple(args)[:2]):12s} pbs {P.n_pbs:6d} levels {P.level_widths}  device {P.last_ms():8.2f} ms  wall {wall:8.2f} ms")
