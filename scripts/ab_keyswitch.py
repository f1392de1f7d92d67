"""A/B of the keyswitch kernels on one GPU: IMAD (0), mma.sync (1), tcgen05.mma kind::i8 (2).  Outputs must be bit-identical; device-timed."""
import sys, json
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import fhe_string_bounty_b200 as F

p = F.Params(**F.PARAM_MESSAGE_2_CARRY_2_KS_PBS)
eng = F.Engine(p)
rng = np.random.default_rng(1)
eng.upload_ksk(rng.integers(0, 2**64, size=p.ksk_len, dtype=np.uint64))
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
rows = []
for B in (37, 128, 1000, 8192):
    d_in = torch.randint(-2**63, 2**63 - 1, (B, p.big_len), dtype=torch.int64, device="cuda")
    ref = None
    for kern in (0, 1, 2):
        eng.set_tuning("ks_kernel", kern)
        d_small = torch.zeros((B, p.small_len), dtype=torch.int64, device="cuda")
        for _ in range(2):
            eng.keyswitch_batch_device(d_in, d_small, B, ts.cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            eng.keyswitch_batch_device(d_in, d_small, B, ts.cuda_stream)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        same = True if ref is None else bool(torch.equal(ref, d_small))
        if ref is None: ref = d_small.clone()
        rows.append(dict(batch=B, ks_kernel=kern, ms=round(ms, 4), identical_to_imad=same))
        print(rows[-1], flush=True)
print(json.dumps(rows))
assert all(r["identical_to_imad"] for r in rows)
