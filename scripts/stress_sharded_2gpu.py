"""2-GPU stress of the sharded string operations (bench.py's config 3 operands): repeated contains / find / case with decryption checks,
interleaved the way bench.py interleaves them.  usage: stress_sharded_2gpu.py [reps] [peer|nccl]"""
import os, sys, socket
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
REPS = int(sys.argv[1]) if len(sys.argv) > 1 else 30
EXCH = sys.argv[2] if len(sys.argv) > 2 else "peer"

def worker(rank, world, port):
    sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    import torch, torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    import fhe_string_bounty_b200 as F
    from fhe_string_bounty_b200 import multi_gpu as MG
    from oracle import oracle as O, radix as R
    from helpers import engine_params
    p = O.params("2_2"); ck = O.ClientKey(p, 0x5EED); sk = O.ServerKey(ck, 0x5EEE)
    params = engine_params(p)
    eng = F.Engine(params, device=rank); eng.upload_ksk(sk.ksk); eng.upload_bsk_std(sk.bsk)
    rng = np.random.default_rng(77)
    rand_str = lambda n: bytes(rng.integers(ord("a"), ord("z") + 1, size=n).tolist())
    hs = rand_str(256); ps = hs[201:217]
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).pin_memory().numpy().view(np.uint64)
    hay, pat, nopat = pin(R.encrypt_string(ck, hs)), pin(R.encrypt_string(ck, ps)), pin(R.encrypt_string(ck, b"0123456789ABCDEF"))
    dec = ck.decrypt_message_and_carry
    comm = MG.DeviceComm(eng, exchange=EXCH)
    log = open(ROOT / "gpurun_out" / f"stress_rank{rank}.log", "w")
    def print(*a, **k):
        log.write(" ".join(str(x) for x in a) + "\n"); log.flush()
    def f(q):
        r = MG.sharded_find(comm, params, hay, q, 256, 16)
        return (dec(r[0]), R.decrypt_radix(ck, r[1:]))
    def c(q):
        return dec(MG.sharded_contains(comm, params, hay, q, 256, 16))
    bad = []
    for it in range(REPS):
        got = (c(pat), f(pat), c(nopat), f(nopat), f(pat))
        if got != (1, (1, 201), 0, got[3], (1, 201)) or got[3][0] != 0:
            bad.append((it, got))
    print(rank, EXCH, f"{len(bad)} bad of {REPS} rounds (contains / find, 256-char haystack, operands alternating):", bad[:4])
    comm.close(); eng.close(); dist.barrier(); dist.destroy_process_group()

if __name__ == "__main__":
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(worker, args=(2, port), nprocs=2, join=True)
