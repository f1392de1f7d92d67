"""A/B of the wide classic kernel's instances on one GPU: device-timed KS-PBS steps at batch 8192 (and whole waves only), per tuning."""
import sys, json
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import fhe_string_bounty_b200 as F

p = F.Params(**F.PARAM_MESSAGE_2_CARRY_2_KS_PBS)
eng = F.Engine(p)
rng = np.random.default_rng(1)
eng.upload_ksk(rng.integers(0, 2**64, size=p.ksk_len, dtype=np.uint64))
eng.upload_bsk_std(rng.integers(0, 2**64, size=p.bsk_len, dtype=np.uint64))
eng.upload_luts(rng.integers(0, 2**64, size=(16, p.lut_len), dtype=np.uint64))
sms = torch.cuda.get_device_properties(0).multi_processor_count
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
rows = []
for cts in [4]:
    for B in (8192,):
        d_in = torch.randint(-2**63, 2**63 - 1, (B, p.big_len), dtype=torch.int64, device="cuda")
        d_idx = (torch.arange(B, device="cuda", dtype=torch.int32) % 16).contiguous()
        d_out = torch.empty_like(d_in)
        for _ in range(2):
            eng.ks_pbs_batch_device(d_in, d_idx, d_out, B, ts.cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            eng.ks_pbs_batch_device(d_in, d_idx, d_out, B, ts.cuda_stream)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        ks, pbs = eng.last_kernel_ms()
        rows.append(dict(wide_cts=cts, batch=B, ms=round(ms, 3), pbs_ms=round(pbs, 3), ks_ms=round(ks, 3), kpbs_per_s=round(B / ms, 2),
                         frac=round(B * 194510848.0 / (pbs * 1e-3) / 1e12 / 36.59, 4)))
        print(rows[-1], flush=True)
print(json.dumps(rows))
