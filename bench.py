#!/usr/bin/env python
"""bench.py -- batched KS-PBS throughput on B200 (BASELINE.json configs[1]) and the CPU reference arm.

    python bench.py --gpus N --steps K --warmup W            # this engine (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores

A "step" is one pass of the hot path -- keyswitch + programmable bootstrap with per-ciphertext lookup
table -- over one batch of `--batch` independent LWE ciphertexts per GPU at PARAM_MESSAGE_2_CARRY_2_KS_PBS.
`value` is device-timed whole-job KS-PBS/s with inputs resident in HBM; `e2e` is the same metric through the
host-buffer C-ABI call (tfhe_b200_ks_pbs_batch) with pinned host inputs and outputs, copies inside the
timed region.  Keys and ciphertexts are uniformly random words: the path's work is data independent
(correctness is the job of tests/, which use real seeded keys against the oracle).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

FLOP_PER_PBS = 742 * 262144.0          # n * ((k+1)(l+1)(5 M log2 M + 6 M) + (k+1)^2 l M 8), M = 1024
BSK_BYTES = 742 * 4 * 1024 * 16        # Fourier bootstrapping key
KSK_BYTES = 2048 * 5 * 743 * 8
CT_BYTES = 2049 * 8


def _clock_sampler(stop, samples, device_index):
    q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    while not stop.is_set():
        try:
            out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(device_index)],
                                 capture_output=True, text=True, timeout=5).stdout.strip()
            if out:
                samples.append([s.strip() for s in out.split(",")])
        except Exception:
            pass
        stop.wait(0.2)


def _summarise_clocks(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
    sm = sorted(float(s[0]) for s in samples if s[0].replace(".", "").isdigit())
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = [n for i, n in enumerate(names) if any(len(s) > 3 + i and s[3 + i].lower().startswith("active") for s in samples)]
    return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(samples[0][1]) if samples[0][1].replace(".", "").isdigit() else None,
            "power_w_max": max((float(s[2]) for s in samples if s[2].replace(".", "").isdigit()), default=None),
            "samples": len(samples), "reasons": reasons}


def _peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        return json.loads(f.read_text()), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class CpuReference:
    """The CPU oracle (restatement of the reference's KS-PBS, one ciphertext per OpenMP thread, the structure of
    tfhe/benches/core_crypto/pbs_bench.rs:512-536) set up once and timed per call.  Only this leg may touch oracle/."""

    def __init__(self, max_cts: int, threads: int = 0, params_name: str = "2_2"):
        from oracle import oracle as O
        import ctypes as C
        self.C, self.L = C, O.lib()
        self.p = O.params(params_name)
        rng = np.random.default_rng(0xB200)
        # random words instead of generated keys: the arithmetic does not depend on the key values
        self.ksk = rng.integers(0, 2**64, size=self.L.orc_ksk_len(self.p), dtype=np.uint64)
        bsk = rng.integers(0, 2**64, size=self.L.orc_bsk_len(self.p), dtype=np.uint64)
        self.f = self.L.orc_fourier_bsk_new(C.byref(self.p), bsk)
        self.luts = rng.integers(0, 2**64, size=(16, self.p.lut_len), dtype=np.uint64)
        self.cts = rng.integers(0, 2**64, size=(max_cts, self.p.big_dim + 1), dtype=np.uint64)
        self.idx = (np.arange(max_cts) % 16).astype(np.uint32)
        self.out = np.zeros_like(self.cts)
        self.cores = threads or self.L.orc_max_threads()

    def run(self, n_cts: int):
        C = self.C
        t0 = time.perf_counter()
        used = self.L.orc_ks_pbs_batch(C.byref(self.p), self.ksk, self.f, self.luts, self.idx.ctypes.data_as(C.c_void_p),
                                       self.cts[:n_cts], self.out[:n_cts], None, n_cts, self.cores)
        return time.perf_counter() - t0, used

    def close(self):
        self.L.orc_fourier_bsk_free(self.f)


def cpu_reference_rate(n_cts: int, threads: int = 0):
    ref = CpuReference(n_cts, threads)
    ref.run(min(n_cts, ref.cores))          # warm-up: page in the Fourier key, spin up the OpenMP team
    dt, used = ref.run(n_cts)
    ref.close()
    return n_cts / dt, used, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ref = CpuReference(1)
    per_step = ref.cores * 32               # about 0.8 s of work per step on the box's 16 cores
    ref.close()
    ref = CpuReference(per_step)
    total_t, used = 0.0, ref.cores
    for s in range(args.warmup + args.steps):
        dt, used = ref.run(per_step)
        if s >= args.warmup:
            total_t += dt
    ref.close()
    value = per_step * args.steps / total_t
    line = {
        "impl": "reference", "metric": "batched KS-PBS throughput", "value": value, "unit": "PBS/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "PARAM_MESSAGE_2_CARRY_2_KS_PBS batched KS-PBS (configs[1])", "batch_per_step": per_step,
                   "note": "reference Rust crate cannot be built here (no cargo); CPU oracle port of its algorithm, all host cores"},
        "cpu_baseline": {"value": value, "unit": "PBS/s", "cores": used, "kind": "port",
                         "sample": f"{per_step} KS-PBS per step x {args.steps} steps, one ciphertext per OpenMP thread"},
        "e2e": {"value": value, "unit": "PBS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_param_sweep(args):
    """--param-sweep: KS-PBS/s of other classic parameter sets (pbs_generic.cu) on one GPU next to the CPU oracle on the
    same box.  Device-resident inputs, CUDA events on the launching stream, random key words (the arithmetic does not depend on the key
    values).  One JSON line; not the headline metric."""
    rows = param_sweep_rows(args.param_sweep.split(","), args.cpu_sample > 0, 0)
    print(json.dumps({"metric": "KS-PBS throughput per classic parameter set (1 GPU, device-resident)", "unit": "PBS/s", "sets": rows}))


def param_sweep_rows(names, with_cpu: bool, device: int):
    import math
    import torch
    import fhe_string_bounty_b200 as F
    torch.cuda.set_device(device)
    sms = torch.cuda.get_device_properties(device).multi_processor_count
    rows = []
    for name in names:
        p = F.Params(**F.classic_params(name))
        eng = F.Engine(p, device=device)
        rng = np.random.default_rng(0xB200)
        eng.upload_ksk(rng.integers(0, 2**64, size=p.ksk_len, dtype=np.uint64))
        eng.upload_bsk_std(rng.integers(0, 2**64, size=p.bsk_len, dtype=np.uint64))
        eng.upload_luts(rng.integers(0, 2**64, size=(4, p.lut_len), dtype=np.uint64))
        small_n = p.poly_size <= 8192
        B = (8 if small_n else 2) * sms
        d_in = torch.randint(-2**63, 2**63 - 1, (B, p.big_len), dtype=torch.int64, device="cuda")
        d_idx = (torch.arange(B, device="cuda", dtype=torch.int32) % 4).contiguous()
        d_out = torch.empty_like(d_in)
        ts = torch.cuda.Stream()
        torch.cuda.set_stream(ts)
        eng.ks_pbs_batch_device(d_in, d_idx, d_out, B, ts.cuda_stream)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3 if small_n else 1
        ev0.record()
        for _ in range(reps):
            eng.ks_pbs_batch_device(d_in, d_idx, d_out, B, ts.cuda_stream)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / reps
        ks_ms, pbs_ms = eng.last_kernel_ms()
        eng.close()
        del d_in, d_out
        torch.cuda.empty_cache()
        M, k1, l = p.poly_size // 2, p.glwe_dim + 1, p.pbs_level
        flop = p.lwe_dim * (k1 * (l + 1) * (5 * M * math.log2(M) + 6 * M) + k1 * k1 * l * M * 8)
        row = {"params": f"PARAM_MESSAGE_{name.replace('_', '_CARRY_')}_KS_PBS", "N": p.poly_size, "k": p.glwe_dim, "pbs_level": l,
               "batch": B, "ms": ms, "keyswitch_ms": ks_ms, "pbs_ms": pbs_ms, "ks_pbs_per_s": B / (ms * 1e-3), "flop_per_pbs": flop,
               "pbs_tflops": B * flop / (pbs_ms * 1e-3) / 1e12}
        if with_cpu:
            ref = CpuReference(1, params_name=name)
            n_cpu = ref.cores * (2 if small_n else 1)
            ref.close()
            ref = CpuReference(n_cpu, params_name=name)
            ref.run(min(n_cpu, ref.cores))
            dt, used = ref.run(n_cpu)
            ref.close()
            row["cpu_port_ks_pbs_per_s"] = n_cpu / dt
            row["cpu_cores"] = used
            row["gpu_over_cpu"] = row["ks_pbs_per_s"] / row["cpu_port_ks_pbs_per_s"]
        rows.append(row)
    return rows


def bench_string_ops(eng, p, rank, world, local):
    """FheString eq / contains / find ops per second (BASELINE.json configs[0], [2]) through the host layer: host buffers in,
    host buffers out (H2D, every tree level's launches, D2H inside the timed region).  With several ranks, eq shards the
    chars and contains shards the 241 windows; the per-rank boolean blocks meet in one NCCL all-reduce (16 KiB)."""
    import torch
    import torch.distributed as dist
    import fhe_string_bounty_b200 as F
    from fhe_string_bounty_b200 import multi_gpu as MG
    from fhe_string_bounty_b200.host import Program
    params = dict(F.PARAM_MESSAGE_2_CARRY_2_KS_PBS)
    rng = np.random.default_rng(77)              # same on every rank (the operands are replicated, the work is sharded)
    execute = lambda prog, ins: prog.run(eng, ins)
    dev = f"cuda:{local}"
    out = {}

    def timed(fn, reps):
        fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t[0])
        return reps / dt

    _pinned = []

    def pin(arr):
        """page-locked copy of a host array (the e2e contract: inputs leave from pinned host memory)"""
        t = torch.from_numpy(arr.view(np.int64)).pin_memory()
        _pinned.append(t)
        return t.numpy().view(np.uint64)

    def run_pinned(prog, ins):
        """single-GPU program through host buffers, both of them page-locked"""
        key = id(prog)
        if key not in run_pinned.out:
            run_pinned.out[key] = pin(np.zeros((prog.n_outputs, p.big_len), dtype=np.uint64))
        return prog.run(eng, ins, out=run_pinned.out[key])
    run_pinned.out = {}

    a8 = pin(rng.integers(0, 2**64, size=(32, p.big_len), dtype=np.uint64))
    b8 = pin(rng.integers(0, 2**64, size=(32, p.big_len), dtype=np.uint64))
    hay = pin(rng.integers(0, 2**64, size=(1024, p.big_len), dtype=np.uint64))
    pat = pin(rng.integers(0, 2**64, size=(64, p.big_len), dtype=np.uint64))
    out["eq_8char_ops_per_s"] = timed(lambda: MG.sharded_eq(execute, params, a8, b8, 8, rank, world, dev), 20)
    out["contains_256_16_ops_per_s"] = timed(lambda: MG.sharded_contains(execute, params, hay, pat, 256, 16, rank, world, dev), 3)
    # throughput mode: independent string pairs share each tree level's launches (each rank takes its own share of pairs)
    n_pairs = 512
    many = Program("string_eq_many", (8, 8, n_pairs), params=params)
    pairs = pin(rng.integers(0, 2**64, size=(many.n_inputs, p.big_len), dtype=np.uint64))
    out["eq_8char_batched_ops_per_s"] = world * n_pairs * timed(lambda: run_pinned(many, pairs), 3)
    out["eq_8char_batch"] = {"pairs_per_rank": n_pairs, "pbs": many.n_pbs, "levels": many.level_widths}
    manyp = Program("string_eq_many_packed", (8, 8, n_pairs), params=params)    # one PBS per pair of blocks (comparator.rs:193-221)
    out["eq_8char_batched_packed_ops_per_s"] = world * n_pairs * timed(lambda: run_pinned(manyp, pairs), 3)
    out["eq_8char_batch_packed"] = {"pairs_per_rank": n_pairs, "pbs": manyp.n_pbs, "levels": manyp.level_widths}
    if world > 1:
        # the other shardings of SURVEY 8(e): find = windows split + all-gather of (found, index) + first-rank selection;
        # to_lowercase = chars split + all-gather of the converted blocks
        # (every rank takes the same path, so an error here is the same on all of them and cannot strand a collective; it must not cost
        # the headline line)
        try:
            out["find_256_16_ops_per_s"] = timed(lambda: MG.sharded_find(execute, params, hay, pat, 256, 16, rank, world, dev), 3)
        except Exception as e:
            out["find_256_16_error"] = str(e)[:200]
        try:
            s1024 = pin(rng.integers(0, 2**64, size=(4096, p.big_len), dtype=np.uint64))
            out["to_lowercase_1024_ops_per_s"] = timed(lambda: MG.sharded_case(execute, params, "to_lowercase", s1024, 1024, rank, world, dev, gather=False), 3)
        except Exception as e:
            out["to_lowercase_1024_error"] = str(e)[:200]
    if world == 1:
        cp = Program("string_contains_packed", (256, 16), params=params)
        ins_c = pin(np.concatenate([hay, pat]))
        out["contains_256_16_packed_ops_per_s"] = timed(lambda: run_pinned(cp, ins_c), 3)
        out["contains_256_16_packed_pbs"] = cp.n_pbs
        find = Program("string_find", (256, 16), params=params)
        ins = ins_c
        out["find_256_16_ops_per_s"] = timed(lambda: run_pinned(find, ins), 3)
        out["find_256_16_pbs"] = find.n_pbs
        findp = Program("string_find_packed", (256, 16), params=params)
        out["find_256_16_packed_ops_per_s"] = timed(lambda: run_pinned(findp, ins), 3)
        out["find_256_16_packed_pbs"] = findp.n_pbs
        low = Program("string_to_lowercase", (1024,), params=params)
        s1024 = pin(rng.integers(0, 2**64, size=(4096, p.big_len), dtype=np.uint64))
        out["to_lowercase_1024_ops_per_s"] = timed(lambda: run_pinned(low, s1024), 3)
        out["to_lowercase_1024_pbs"] = low.n_pbs
        up = Program("string_to_uppercase", (1024,), params=params)
        out["to_uppercase_1024_ops_per_s"] = timed(lambda: run_pinned(up, s1024), 3)
        eic = Program("string_eq_ignore_case", (1024, 1024), params=params)
        two = pin(rng.integers(0, 2**64, size=(eic.n_inputs, p.big_len), dtype=np.uint64))
        out["eq_ignore_case_1024_ops_per_s"] = timed(lambda: run_pinned(eic, two), 2)
        out["eq_ignore_case_1024_pbs"] = eic.n_pbs
    out["note"] = ("page-locked host buffers in/out; eq shards chars, contains shards windows across ranks + one all-reduce of a 2049-word LWE; "
                   "with several ranks find shards windows (all-gather of found + index, 2 selection levels) and to_lowercase shards chars (no exchange: each rank keeps its converted chars)")
    return out


def bench_string_ops_multibit(eng, p, rank, world):
    """BASELINE.json configs[4]: lexicographic lt / le on 128-char strings with the multi-bit parameter set.  A comparison is one
    10-level tree ([256, 128, ..., 1, 1] blocks) that does not shard, so every rank runs whole comparisons (replicas): ops/s = ranks x
    one rank's rate, latency = one rank's time."""
    import torch
    import torch.distributed as dist
    import fhe_string_bounty_b200 as F
    from fhe_string_bounty_b200.host import Program
    params = dict(F.PARAM_MULTI_BIT_MESSAGE_2_CARRY_2_GROUP_3_KS_PBS)
    rng = np.random.default_rng(78 + rank)
    out = {}
    for op in ("string_lt", "string_le"):
        prog = Program(op, (128, 128), params=params)
        ins = rng.integers(0, 2**64, size=(prog.n_inputs, p.big_len), dtype=np.uint64)
        prog.run(eng, ins)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            prog.run(eng, ins)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t[0])
        out[op[7:] + "_128char_ms"] = 1e3 * dt / reps
        out[op[7:] + "_128char_ops_per_s"] = world * reps / dt
        out[op[7:] + "_128char_pbs"] = prog.n_pbs
        out[op[7:] + "_128char_levels"] = prog.level_widths
    out["note"] = "host buffers in/out; a comparison tree does not shard: replicas only (each rank compares its own pair of strings)"
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist
    import fhe_string_bounty_b200 as F

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL writes its version banner (any NCCL_DEBUG level from VERSION up, which the GPU boxes set) and its warnings to stdout by
        # default: send them to stderr so that stdout stays the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    B = args.batch
    p = F.Params(**(F.PARAM_MULTI_BIT_MESSAGE_2_CARRY_2_GROUP_3_KS_PBS if args.params == "multibit" else F.PARAM_MESSAGE_2_CARRY_2_KS_PBS))
    eng = F.Engine(p, device=local)
    rng = np.random.default_rng(0xB200)          # same keys on every rank: keys are replicated, work is sharded
    eng.upload_ksk(rng.integers(0, 2**64, size=p.ksk_len, dtype=np.uint64))
    eng.upload_bsk_std(rng.integers(0, 2**64, size=p.bsk_len, dtype=np.uint64))
    eng.upload_luts(rng.integers(0, 2**64, size=(16, p.lut_len), dtype=np.uint64))

    g = torch.Generator(device="cuda").manual_seed(1234 + rank)
    d_in = torch.randint(-2**63, 2**63 - 1, (B, p.big_len), dtype=torch.int64, device="cuda", generator=g)
    d_idx = (torch.arange(B, device="cuda", dtype=torch.int32) % 16).contiguous()
    d_out = torch.empty_like(d_in)
    # a non-default torch stream: handle 0 would mean "the context's own stream" to the C ABI, and
    # torch.cuda.Event only sees the stream it is recorded on
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    def step():
        eng.ks_pbs_batch_device(d_in, d_idx, d_out, B, stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    stop, samples = threading.Event(), []
    sampler = threading.Thread(target=_clock_sampler, args=(stop, samples, local), daemon=True)
    sampler.start()
    launches0 = eng.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ks_ms_sum = pbs_ms_sum = 0.0
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = eng.kernel_launches - launches0
    # per-kernel durations of the dominant kernel, measured live (CUDA events on the launching stream)
    for _ in range(min(args.steps, 5)):
        step()
        torch.cuda.synchronize()
        k, q = eng.last_kernel_ms()
        ks_ms_sum += k
        pbs_ms_sum += q
    n_k = min(args.steps, 5)
    ks_ms, pbs_ms = ks_ms_sum / n_k, pbs_ms_sum / n_k
    stop.set()
    sampler.join(timeout=2)

    # e2e through the host-buffer C-ABI entry point, pinned host memory, copies inside the timed region
    h_in = torch.empty((B, p.big_len), dtype=torch.int64).pin_memory()
    h_in.copy_(d_in.cpu())
    h_idx = (torch.arange(B, dtype=torch.int32) % 16).pin_memory()
    h_out = torch.empty((B, p.big_len), dtype=torch.int64).pin_memory()
    e2e_steps = max(2, min(args.steps, 5))
    eng.ks_pbs_batch(h_in, h_idx, h_out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        eng.ks_pbs_batch(h_in, h_idx, h_out)
    barrier()
    e2e_s = time.perf_counter() - t0

    string_ops = None
    if args.string_ops and args.params == "2_2":
        string_ops = bench_string_ops(eng, p, rank, world, local)
    elif args.string_ops:
        string_ops = bench_string_ops_multibit(eng, p, rank, world)

    if world > 1:
        t = torch.tensor([ms, e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(t[0]), float(t[1])
    if rank == 0:
        peaks, peak_kind = _peaks()
        fp64_peak = eng.probe_fp64_tflops()
        value = B * world * args.steps / (ms * 1e-3)
        e2e = B * world * e2e_steps / e2e_s
        pbs_tflops = B * FLOP_PER_PBS / (pbs_ms * 1e-3) / 1e12
        hbm_alg = (BSK_BYTES + KSK_BYTES + B * (2 * CT_BYTES + 743 * 8 * 2)) / ((pbs_ms + ks_ms) * 1e-3) / 1e9
        cpu_rate, cpu_cores, cpu_dt = cpu_reference_rate(args.cpu_sample) if args.cpu_sample > 0 else (None, 0, 0)
        traffic = None
        tf = ROOT / "profiles" / "r01_pbs_traffic.json"
        if tf.exists() and args.params == "2_2":
            t = json.loads(tf.read_text())
            if t.get("batch") == B:        # dram__bytes_read+write of one launch from the committed ncu capture of this workload
                traffic = t["dram_bytes"]
        line = {
            "metric": "batched KS-PBS throughput", "value": value, "unit": "PBS/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": ("PARAM_MESSAGE_2_CARRY_2_KS_PBS" if args.params == "2_2" else "PARAM_MULTI_BIT_MESSAGE_2_CARRY_2_GROUP_3_KS_PBS") + " batched KS-PBS (configs[1])", "batch_per_gpu": B,
                       "luts": 16, "l2_policy": "inputs larger than L2 (cts in+out %.0f MB + keys 109 MB per step)" % (2 * B * CT_BYTES / 1e6),
                       "sharding": "ciphertexts partitioned across ranks, keys replicated, no data-path collective"},
            "e2e": {"value": e2e, "unit": "PBS/s", "h2d_bytes_per_step": B * (CT_BYTES + 4), "d2h_bytes_per_step": B * CT_BYTES},
            "gpu_launches": launches,
            "kernels": {"keyswitch_ms": ks_ms, "pbs_ms": pbs_ms},
            "roofline": {"bound": "fp64", "kernel": "pbs_classic_kernel_v4" if args.params == "2_2" else "pbs_multibit_kernel", "achieved": pbs_tflops, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": pbs_tflops / fp64_peak if fp64_peak else None, "traffic": traffic,
                         "algorithmic_bytes_per_launch": BSK_BYTES + B * (743 * 8 + CT_BYTES) + 16 * 4096 * 8,
                         "peak_source": "FP64 FMA microbenchmark run in this process (MEASURED_PEAKS.json has no FP64 figure)",
                         "flop_per_pbs": FLOP_PER_PBS,
                         "hbm": {"achieved": hbm_alg, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                                 "frac": hbm_alg / peaks["hbm_gbs"] if peaks.get("hbm_gbs") else None, "peak_source": peak_kind}},
            "clocks": _summarise_clocks(samples),
        }
        if string_ops is not None:
            line["string_ops"] = string_ops
        if args.other_sets and world == 1:
            # breadth, next to the headline: two other classic parameter sets on the generic kernel, this rank's GPU only
            try:
                line["other_parameter_sets"] = [
                    {k: r[k] for k in ("params", "N", "k", "pbs_level", "batch", "ks_pbs_per_s", "pbs_ms", "pbs_tflops")}
                    for r in param_sweep_rows(args.other_sets.split(","), False, local)]
            except Exception as e:            # never lose the headline over the extra rows
                line["other_parameter_sets"] = {"error": str(e)[:200]}
        if cpu_rate is not None and string_ops is not None:
            # the CPU path has no batching effect beyond its cores: a string op costs (its PBS count) / (CPU KS-PBS rate)
            string_ops["cpu_port_ops_per_s_derived"] = {
                "eq_8char": cpu_rate / 36.0, "contains_256_16": cpu_rate / 16890.0, "find_256_16": cpu_rate / 17916.0,
                "to_lowercase_1024": cpu_rate / 4096.0, "note": "derived: CPU KS-PBS/s of cpu_baseline / PBS count of the reference-shaped tree"}
        if cpu_rate is not None:
            line["cpu_baseline"] = {"value": cpu_rate, "unit": "PBS/s", "cores": cpu_cores, "kind": "port",
                                    "sample": f"{args.cpu_sample} KS-PBS (same parameter set, random keys), {cpu_dt:.1f} s"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    eng.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=8192, help="ciphertexts per GPU per step")
    ap.add_argument("--params", default="2_2", choices=["2_2", "multibit"], help="2_2 = PARAM_MESSAGE_2_CARRY_2_KS_PBS (headline); multibit = ..._GROUP_3_KS_PBS")
    ap.add_argument("--string-ops", type=int, default=1, help="also time FheString eq/contains/find through the host layer (0 = skip)")
    ap.add_argument("--param-sweep", default="", help="comma-separated classic sets (e.g. 1_1,3_3,4_4): per-set KS-PBS/s on one GPU + CPU oracle, then exit")
    ap.add_argument("--other-sets", default="1_1,3_3", help="classic sets measured briefly after the headline and reported as other_parameter_sets ('' = skip)")
    ap.add_argument("--cpu-sample", type=int, default=8192, help="KS-PBS evaluated by the CPU baseline leg (0 = skip)")
    args = ap.parse_args()
    if args.param_sweep:
        run_param_sweep(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
