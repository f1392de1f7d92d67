import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import fhe_string_bounty_b200 as F
p = F.Params(**F.PARAM_MESSAGE_2_CARRY_2_KS_PBS)
eng = F.Engine(p)
rng = np.random.default_rng(1)
eng.upload_ksk(rng.integers(0, 2**64, size=p.ksk_len, dtype=np.uint64))
eng.upload_bsk_std(rng.integers(0, 2**64, size=p.bsk_len, dtype=np.uint64))
eng.upload_luts(rng.integers(0, 2**64, size=(4, p.lut_len), dtype=np.uint64))
s = torch.cuda.Stream()
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 148
d_in = torch.from_numpy(rng.integers(0, 2**63, size=(batch, p.big_len), dtype=np.int64)).cuda()
d_out = torch.empty_like(d_in)
idx = torch.zeros(batch, dtype=torch.int32, device="cuda")
for _ in range(4):
    eng.ks_pbs_batch_device(d_in, idx, d_out, batch, s.cuda_stream)
s.synchronize()
print("done")
