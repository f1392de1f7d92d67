"""PARAM_MESSAGE_1_CARRY_1_KS_PBS: device-timed KS-PBS on the tuned kernel (pbs_n512.cu) against the generic kernel."""
import sys, json
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import fhe_string_bounty_b200 as F

p = F.Params(**F.classic_params("1_1"))
eng = F.Engine(p)
rng = np.random.default_rng(1)
eng.upload_ksk(rng.integers(0, 2**64, size=p.ksk_len, dtype=np.uint64))
eng.upload_bsk_std(rng.integers(0, 2**64, size=p.bsk_len, dtype=np.uint64))
eng.upload_luts(rng.integers(0, 2**64, size=(4, p.lut_len), dtype=np.uint64))
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
FLOP = 86851584.0
for B in (1184, 8192, 8 * 148 * 8):
    d_in = torch.randint(-2**63, 2**63 - 1, (B, p.big_len), dtype=torch.int64, device="cuda")
    d_idx = (torch.arange(B, device="cuda", dtype=torch.int32) % 4).contiguous()
    d_out = torch.empty_like(d_in)
    for name, mn in (("tuned", 1), ("generic", 1 << 30)):
        eng.set_tuning("tuned512_min", mn)
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
        print(dict(kernel=name, batch=B, ms=round(ms, 3), pbs_ms=round(pbs, 3), ks_ms=round(ks, 3), kpbs_per_s=round(B / ms, 2),
                   pbs_tflops=round(B * FLOP / (pbs * 1e-3) / 1e12, 2)), flush=True)
