"""2-GPU debug: per-rank share of sharded_find through the normal output path and through the exchange send area, both exchanges."""
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
    p = O.params("2_2"); ck = O.ClientKey(p, 0xB200 + 1); sk = O.ServerKey(ck, 0xB300 + 1)
    params = engine_params(p)
    eng = F.Engine(params, device=rank); eng.upload_ksk(sk.ksk); eng.upload_bsk_std(sk.bsk)
    rng = np.random.default_rng(0xB200 + 3)
    hay = bytes(rng.integers(ord("a"), ord("z") + 1, size=64).tolist()); pat = hay[37:45]
    h, q = R.encrypt_string(ck, hay), R.encrypt_string(ck, pat)
    dec = lambda rows: [ck.decrypt_message_and_carry(r) for r in rows]
    comm = MG.DeviceComm(eng, exchange="peer")
    w0, w1 = MG.shard_range(57, rank, 2)
    prog = comm.program("string_find_windows", (64, 8, w0, w1), params)
    plain = comm.to_host(comm.run(prog, [h, q]))
    print(rank, "share via normal output:", dec(plain), flush=True)
    for it in range(3):
        comm.run(prog, [h, q], to_send_area=True)
        parts = comm.to_host(comm.all_gather(None).reshape(-1, eng.p.big_len))
        print(rank, f"gather #{it} via send area:", dec(parts), flush=True)
    parts = comm.to_host(comm.all_gather(comm._to_device(plain)).reshape(-1, eng.p.big_len))
    print(rank, "gather of staged rows:", dec(parts), flush=True)
    out = MG.sharded_find(comm, params, h, q, 64, 8)
    print(rank, "sharded_find peer:", dec(out), flush=True)
    out = MG.sharded_contains(comm, params, h, q, 64, 8)
    print(rank, "sharded_contains peer:", dec([out]), flush=True)
    out = MG.sharded_find(comm, params, h, q, 64, 8)
    print(rank, "sharded_find peer again:", dec(out), flush=True)
    comm.close()
    comm = MG.DeviceComm(eng, exchange="nccl")
    out = MG.sharded_find(comm, params, h, q, 64, 8)
    print(rank, "sharded_find nccl:", dec(out), flush=True)
    comm.close(); eng.close(); dist.barrier(); dist.destroy_process_group()

if __name__ == "__main__":
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(worker, args=(2, port), nprocs=2, join=True)
