"""One parameter set, one batch size, a few KS-PBS launches on random keys: the workload ncu captures of the non-headline kernels run
(usage: python scripts/profile_case.py <classic "m_c" | multi-bit "m_c_gG"> <batch>)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import fhe_string_bounty_b200 as F

name, batch = sys.argv[1], int(sys.argv[2])
p = F.Params(**(F.multi_bit_params(name) if "_g" in name else F.classic_params(name)))
eng = F.Engine(p)
rng = np.random.default_rng(1)
eng.upload_ksk(rng.integers(0, 2**64, size=p.ksk_len, dtype=np.uint64))
eng.upload_bsk_std(rng.integers(0, 2**64, size=p.bsk_len, dtype=np.uint64))
eng.upload_luts(rng.integers(0, 2**64, size=(4, p.lut_len), dtype=np.uint64))
s = torch.cuda.Stream()
d_in = torch.from_numpy(rng.integers(0, 2**63, size=(batch, p.big_len), dtype=np.int64)).cuda()
d_out = torch.empty_like(d_in)
idx = torch.zeros(batch, dtype=torch.int32, device="cuda")
for _ in range(3):
    eng.ks_pbs_batch_device(d_in, idx, d_out, batch, s.cuda_stream)
s.synchronize()
print("done", eng.last_kernel_ms())
eng.close()
