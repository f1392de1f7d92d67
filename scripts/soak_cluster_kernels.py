"""Soak of the two-SM cluster kernels (st.async / mbarrier hand-offs): many launches of narrow levels of several widths on the classic,
the multi-bit and the 3_3 set; every run must return the words of the first run.  usage: soak_cluster_kernels.py [reps]"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import fhe_string_bounty_b200 as F

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
sets = [("2_2", F.PARAM_MESSAGE_2_CARRY_2_KS_PBS, (1, 2, 37, 73, 74)), ("2_2 multi-bit g3", F.PARAM_MULTI_BIT_MESSAGE_2_CARRY_2_GROUP_3_KS_PBS, (1, 2, 37, 73, 74)),
        ("3_3", F.classic_params("3_3"), (1, 74, 75, 150))]
bad_total = 0
for name, params, widths in sets:
    p = F.Params(**params)
    eng = F.Engine(p)
    rng = np.random.default_rng(7)
    eng.upload_ksk(rng.integers(0, 2**64, size=p.ksk_len, dtype=np.uint64))
    eng.upload_bsk_std(rng.integers(0, 2**64, size=p.bsk_len, dtype=np.uint64))
    eng.upload_luts(rng.integers(0, 2**64, size=(4, p.lut_len), dtype=np.uint64))
    s = torch.cuda.Stream()
    for batch in widths:
        n = max(4, reps // 20) if name == "3_3" else reps
        d_in = torch.from_numpy(rng.integers(0, 2**63, size=(batch, p.big_len), dtype=np.int64)).cuda()
        idx = (torch.arange(batch, dtype=torch.int32, device="cuda") % 4).contiguous()
        ref = torch.empty_like(d_in); out = torch.empty_like(d_in)
        t0 = time.time()
        with torch.cuda.stream(s):
            eng.ks_pbs_batch_device(d_in, idx, ref, batch, s.cuda_stream)
            bad = 0
            for r in range(n):
                eng.ks_pbs_batch_device(d_in, idx, out, batch, s.cuda_stream)
                if r % 10 == 9 or r == n - 1:
                    s.synchronize()
                    bad += int(not torch.equal(out, ref))
        s.synchronize()
        bad_total += bad
        print(f"{name:18s} width {batch:4d}: {n} launches, {bad} differing checks, {time.time() - t0:.1f} s", flush=True)
    eng.close()
print("TOTAL differing checks:", bad_total)
sys.exit(1 if bad_total else 0)
