"""Device-timed KS-PBS of a tuned kernel against the generic kernel on the same context.  usage: ab_tuned.py 1_1|3_3 [batches...]"""
import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import fhe_string_bounty_b200 as F

name = sys.argv[1]
batches = [int(x) for x in sys.argv[2:]] or [1184]
key, on, off = {"1_1": ("tuned512_min", 1, 1 << 30), "3_3": ("tuned8192", 1, 0)}[name]
p = F.Params(**F.classic_params(name))
eng = F.Engine(p)
rng = np.random.default_rng(1)
eng.upload_ksk(rng.integers(0, 2**64, size=p.ksk_len, dtype=np.uint64))
eng.upload_bsk_std(rng.integers(0, 2**64, size=p.bsk_len, dtype=np.uint64))
eng.upload_luts(rng.integers(0, 2**64, size=(4, p.lut_len), dtype=np.uint64))
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
k1, N, n, l = p.glwe_dim + 1, p.poly_size, p.lwe_dim, p.pbs_level
M = N // 2
FLOP = n * ((k1 * l + k1) * 5 * M * np.log2(M) + k1 * l * k1 * M * 8)      # forward + inverse transforms + multiply-accumulate
for B in batches:
    d_in = torch.randint(-2**63, 2**63 - 1, (B, p.big_len), dtype=torch.int64, device="cuda")
    d_idx = (torch.arange(B, device="cuda", dtype=torch.int32) % 4).contiguous()
    d_out = torch.empty_like(d_in)
    for label, v in (("tuned", on), ("generic", off)):
        eng.set_tuning(key, v)
        for _ in range(2):
            eng.ks_pbs_batch_device(d_in, d_idx, d_out, B, ts.cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2):
            eng.ks_pbs_batch_device(d_in, d_idx, d_out, B, ts.cuda_stream)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 2
        ks, pbs = eng.last_kernel_ms()
        print(dict(set=name, kernel=label, batch=B, ms=round(ms, 3), pbs_ms=round(pbs, 3), ks_ms=round(ks, 3), pbs_per_s=round(B / ms * 1e3, 1),
                   pbs_tflops=round(B * FLOP / (pbs * 1e-3) / 1e12, 2)), flush=True)
