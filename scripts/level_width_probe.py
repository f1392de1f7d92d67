import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import fhe_string_bounty_b200 as F
MB = "--mb" in sys.argv        # multi-bit GROUP_3 set instead of the classic headline set
if MB: sys.argv.remove("--mb")
p = F.Params(**(F.PARAM_MULTI_BIT_MESSAGE_2_CARRY_2_GROUP_3_KS_PBS if MB else F.PARAM_MESSAGE_2_CARRY_2_KS_PBS))
eng = F.Engine(p)
rng = np.random.default_rng(1)
eng.upload_ksk(rng.integers(0, 2**64, size=p.ksk_len, dtype=np.uint64))
eng.upload_bsk_std(rng.integers(0, 2**64, size=p.bsk_len, dtype=np.uint64))
eng.upload_luts(rng.integers(0, 2**64, size=(4, p.lut_len), dtype=np.uint64))
s = torch.cuda.Stream()
for batch in [int(a) for a in sys.argv[1:]] or (74, 148, 296, 297, 400, 444, 445, 592, 700, 888, 1024, 1036, 1184):
    d_in = torch.from_numpy(rng.integers(0, 2**63, size=(batch, p.big_len), dtype=np.int64)).cuda()
    d_out = torch.empty_like(d_in)
    idx = torch.zeros(batch, dtype=torch.int32, device="cuda")
    with torch.cuda.stream(s):
        for _ in range(2):
            eng.ks_pbs_batch_device(d_in, idx, d_out, batch, s.cuda_stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(3):
            eng.ks_pbs_batch_device(d_in, idx, d_out, batch, s.cuda_stream)
        e1.record(s)
    s.synchronize()
    print(f"batch {batch:5d}: {e0.elapsed_time(e1) / 3:7.3f} ms")
