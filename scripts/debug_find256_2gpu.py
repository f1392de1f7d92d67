"""2-GPU debug of bench.py's config 3 find (256-char haystack, 16-char pattern from offset 201): sharded result and per-rank shares
under different kernel selections."""
import os, sys, socket
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent

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
    rand_str(8)
    hs = rand_str(256); ps = hs[201:217]
    hay, pat = R.encrypt_string(ck, hs), R.encrypt_string(ck, ps)
    dec = ck.decrypt_message_and_carry
    want = (1, hs.find(ps))
    n_win = 241
    for exch in ("peer", "nccl"):
        comm = MG.DeviceComm(eng, exchange=exch)
        for (ks, cl) in ((2, 1), (2, 0), (1, 1), (1, 0)):
            eng.set_tuning("ks_kernel", ks); eng.set_tuning("narrow_cluster", cl)
            r = MG.sharded_find(comm, params, hay, pat, 256, 16)
            got = (dec(r[0]), R.decrypt_radix(ck, r[1:]))
            w0, w1 = MG.shard_range(n_win, rank, 2)
            share = comm.to_host(comm.run(comm.program("string_find_windows", (256, 16, w0, w1), params), [hay, pat]))
            one = comm.to_host(comm.run(comm.program("string_find", (256, 16), params), [hay, pat]))
            print(rank, exch, f"ks={ks} cluster={cl}: sharded {got} want {want}; my share [{w0},{w1}) -> found {dec(share[0])} idx {R.decrypt_radix(ck, share[1:])};"
                  f" unsharded {(dec(one[0]), R.decrypt_radix(ck, one[1:]))}", flush=True)
        comm.close()
    eng.close(); dist.barrier(); dist.destroy_process_group()

if __name__ == "__main__":
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(worker, args=(2, port), nprocs=2, join=True)
