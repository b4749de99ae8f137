#!/usr/bin/env python3
"""BASELINE.json configs 3 and 5: DISTINCT synthetic Stwo proofs, generated on the GPU by ssym_stwo_prove_batch.

  python bench_configs.py --config 3 [--log-n 16] [--steps 5]                      # 2^16 proofs + corrupted negatives, 1 B200
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
         bench_configs.py --config 5 [--log-n 20] [--steps 3]                      # 2^20 proofs sharded by index over N B200s

config 3: proof i proves the wide-Fibonacci trace of seed i.  Negatives per SURVEY.md section 8d: proof i with i % 8 = k < 6 gets
  corruption class k (0 flipped Merkle sibling bit, 1 FRI witness + 1, 2 OODS composition sample + 1, 3 nonce + 1, 4 last-layer
  coefficient + 1, 5 queried value + p (non-canonical)); i % 8 in {6, 7} stays valid.  Both semantics are timed
  (PROVER_CONSISTENT: valid proofs accept, negatives reject; REF_LITERAL: everything rejects, as the reference's own fixtures do).
config 5: contiguous shards of proof index (sharding.shard_range), each rank proves + verifies its shard in HBM, the accept bitmaps
  are all-gathered (the job's only exchange, inside the timed region), max over ranks of the CUDA-event time.
Every number: inputs resident in HBM and far larger than L2 (>= 3.5 GB per pass), CUDA events on the launching stream, >= 3 warm-up
passes.  A sample of the statuses is compared with the CPU oracle outside the timed region; the CPU baseline is the oracle on all
host cores over a bounded sample of the same proofs.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

import bench as B


def negatives_plan(S, cfg):
    cls = S.witness.stwo_negative_classes(cfg)
    names = ["trace_sibling_bit", "fri_witness_plus_1", "oods_cp_plus_1", "pow_nonce_plus_1", "last_coeff_plus_1", "queried_value_plus_p"]
    return [(k, n, *cls[n]) for k, n in enumerate(names)]


def apply_negatives(torch, proofs, plan, first_index):
    """proof (first_index + r) with (first_index + r) % 8 == k gets class k.  In place, on the device."""
    n = proofs.shape[0]
    idx = torch.arange(first_index, first_index + n, device=proofs.device)
    for k, _, word, delta in plan:
        rows = torch.nonzero(idx % 8 == k).view(-1)
        d = delta if delta < 2**31 else delta - 2**32
        proofs[rows, word] += d  # int32 wrap-around = add mod 2^32
    return (idx % 8 < len(plan)).cpu().numpy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True, choices=[3, 5])
    ap.add_argument("--gpus", type=int, default=None, help="informational (the world size comes from torchrun), accepted for symmetry with bench.py")
    ap.add_argument("--log-n", type=int, default=None)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--out", default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--columns", type=int, default=4, choices=[4, 8, 16], help="NUM_COLUMNS of the AIR (config.simf:14); 4 = reference HEAD")
    args = ap.parse_args()
    # libraries (NCCL's version banner) write to fd 1: keep stdout for the single JSON line, as bench.py does
    sys.stdout.flush()
    B._REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist

    rank, world, local_rank = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    import stark_symphony_b200 as S
    from importlib import import_module

    sharding = import_module("stark_symphony_b200.sharding")
    log_n = args.log_n or (16 if args.config == 3 else 20)
    steps = args.steps or (5 if args.config == 3 else 3)
    n_total = 1 << log_n
    begin, end = sharding.shard_range(n_total, rank, world)
    n = end - begin
    cfg_pc = S.stwo_config("prod", S.MODE_PROVER_CONSISTENT, n_columns=args.columns)
    cfg_lit = S.stwo_config("prod", S.MODE_REF_LITERAL, n_columns=args.columns)
    lo = S.stwo_layout(cfg_pc)
    ver = S.Verifier(local_rank)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ver.set_stream(stream.cuda_stream)

    # ---- generate this rank's proofs on the device ------------------------------------------------------------
    seeds = torch.arange(begin, end, dtype=torch.int64, device="cuda")
    proofs = torch.empty((n, lo.stride_words), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    l0 = ver.launch_count
    t0 = time.perf_counter()
    ver.stwo_prove_batch(seeds, cfg_pc, out=proofs)  # synchronous
    torch.cuda.synchronize()
    prove_s = time.perf_counter() - t0
    prove_launches = ver.launch_count - l0
    B.log(f"rank {rank}: proved {n} proofs in {prove_s:.2f} s ({n / prove_s:.0f} proofs/s, {prove_launches} launches)")
    expect_bad = np.zeros(n, dtype=bool)
    plan = negatives_plan(S, cfg_pc)
    if args.config == 3:
        expect_bad = apply_negatives(torch, proofs, plan, begin)

    words = sharding.shard_words(n_total, world)
    accept = torch.zeros(words, dtype=torch.int32, device="cuda")
    status = torch.zeros(n, dtype=torch.int32, device="cuda")
    gathered = torch.zeros(words * world, dtype=torch.int32, device="cuda")

    def one_pass(cfg):
        ver.stwo_verify_batch(proofs.view(-1), cfg, n, accept_out=accept, status_out=status)
        if world > 1:
            dist.all_gather_into_tensor(gathered, accept)  # the job's only exchange: 4 bytes per 32 proofs

    results = {}
    for name, cfg in (("prover-consistent", cfg_pc), ("ref-literal", cfg_lit)):
        for _ in range(args.warmup):
            one_pass(cfg)
        torch.cuda.synchronize()
        l0 = ver.launch_count
        with B.ClockSampler(local_rank) as clocks:
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(steps):
                one_pass(cfg)
            e1.record(stream)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        st = status.cpu().numpy().view(np.uint32)
        if name == "prover-consistent":
            assert ((st != 0) == expect_bad).all(), f"rank {rank}: {int(((st != 0) != expect_bad).sum())} proofs with the wrong verdict"
        else:
            assert (st != 0).all()
        bits = (gathered if world > 1 else accept).cpu().numpy()
        accepted = sharding.expected_accept_count(bits, words * 32 * world) if world > 1 else sharding.expected_accept_count(bits, n)
        results[name] = {"ms_per_pass": ms / steps, "proofs_per_s": n_total * steps / (ms * 1e-3), "accepted": accepted, "gpu_launches": ver.launch_count - l0,
                         "clocks": clocks.summary(), "status_sample": st[:512].copy()}

    # per-kernel timing (serial, events around every kernel) for the roofline of the dominant kernel
    ver.profile_read()
    ver.profile_enable(True)
    one_pass(cfg_pc)
    ver.profile_enable(False)
    prof = ver.profile_read()
    int32_ops, probe_ms = ver.int32_peak_probe()

    if rank == 0:
        from oracle import oracle as O

        orc = O.Oracle()
        sample_n = min(n, 256)
        sample = proofs[:sample_n].cpu().numpy().view(np.uint32)
        checks = {}
        for name, mode in (("prover-consistent", O.MODE_PROVER_CONSISTENT), ("ref-literal", O.MODE_REF_LITERAL)):
            _, o_status, _ = orc.stwo_verify_batch(O.make_config("prod", mode, args.columns), sample.ravel(), sample_n)
            assert (o_status == results[name]["status_sample"][:sample_n]).all(), f"{name}: GPU statuses differ from the oracle"
            checks[name] = f"{sample_n} statuses bit-identical to the oracle"
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        mk_ms, mk_n = prof.get("stwo_merkle", (0.0, 0))
        kernel_ms = {k: v[0] for k, v in prof.items()}
        # a 16-column trace leaf is a 64-byte message: one more compression per trace decommitment
        merkle_compressions = B.MERKLE_COMPRESSIONS_PER_PROOF + (cfg_pc.n_queries if args.columns == 16 else 0)
        alg_bytes = lo.algorithmic_bytes
        lit_ops = n * merkle_compressions * B.LITERAL_OPS_PER_COMPRESSION / (mk_ms * 1e-3) if mk_ms else 0.0
        r = results["prover-consistent"]
        line = {
            "metric": "stwo_proofs_verified_per_s", "value": r["proofs_per_s"], "unit": "proofs/s", "n_gpus": world, "steps": steps, "warmup": args.warmup,
            "ms_per_step": r["ms_per_pass"], "higher_is_better": True, "scaling": "strong" if args.config == 5 else "weak", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic: distinct proofs of the wide-Fibonacci AIR, one per seed, generated on the GPU by ssym_stwo_prove_batch",
            "config": {"workload": (f"BASELINE config 3: 2^{log_n} distinct synthetic Stwo proofs + corrupted negatives (6 classes x 1/8 of the batch), 1 B200" if args.config == 3 else
                                    f"BASELINE config 5: 2^{log_n} distinct synthetic Stwo proofs sharded by index over {world} B200, accept-bitmap gather"),
                       "mode": "prover-consistent", "n_columns": args.columns, "proofs_total": n_total, "proofs_per_gpu": n, "negatives": int(expect_bad.sum()),
                       "l2": f"one pass reads {n * lo.stride_words * 4 / 1e9:.2f} GB per GPU (>> 126 MB L2)"},
            "accepted": r["accepted"], "gpu_launches": int(r["gpu_launches"]),
            "ref_literal": {"proofs_per_s": results["ref-literal"]["proofs_per_s"], "ms_per_pass": results["ref-literal"]["ms_per_pass"], "accepted": results["ref-literal"]["accepted"]},
            "merkle_hashes_per_s": r["proofs_per_s"] * merkle_compressions / 2.0,
            "kernel_ms_per_pass": kernel_ms,
            "roofline": {"bound": "hbm", "kernel": "stwo_merkle_kernel", "achieved": n * alg_bytes / (mk_ms * 1e-3) / 1e9 if mk_ms else None, "peak": hbm_peak,
                         "unit": "GB/s", "frac": (n * alg_bytes / (mk_ms * 1e-3) / 1e9 / hbm_peak) if mk_ms else None, "traffic": None,
                         "note": "INT32-ALU bound kernel (170 int-ops per byte): see roofline_int32"},
            "roofline_int32": {"bound": "int32_alu", "kernel": "stwo_merkle_kernel", "achieved": lit_ops / 1e12, "peak": int32_ops / 1e12, "unit": "Tops/s",
                               "frac": lit_ops / int32_ops if int32_ops else None, "launches": mk_n, "total_ms": mk_ms},
            "prover": {"proofs_per_s_per_gpu": n / prove_s, "seconds": prove_s, "gpu_launches": int(prove_launches),
                       "note": "ssym_stwo_prove_batch, device resident, wall clock of the synchronous call (workload generation, not the headline metric)"},
            "parity": checks, "clocks": r["clocks"],
        }
        if not args.no_cpu_baseline:
            threads = B.cpu_threads()
            ocfg = O.make_config("prod", O.MODE_PROVER_CONSISTENT, args.columns)
            flat = np.ascontiguousarray(sample.ravel())
            t1, _ = B.oracle_run(orc, ocfg, flat, sample_n, 1, 8)
            total = int(max(threads * 8, min(15.0 / (t1 / 8), 64 * sample_n)))
            dt, done = B.oracle_run(orc, ocfg, flat, sample_n, threads, total)
            line["cpu_baseline"] = {"value": done / dt, "unit": "proofs/s", "cores": threads, "kind": "port",
                                    "sample": f"{done} proofs (cycling the first {sample_n} proofs of the batch) on {threads} threads, {dt:.2f} s wall; oracle/ssym_oracle.c"}
        B.emit(line)
        if args.out:
            with open(args.out, "w") as f:
                json.dump(line, f)
    ver.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
