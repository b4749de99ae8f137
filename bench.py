#!/usr/bin/env python3
"""Headline benchmark: Stwo proofs verified per second on N B200s (BASELINE.json metric).

Workload (BASELINE.json configs[1]): the reference's own `stwo-verifier/tests/data/proof.json` witness (prod preset:
LDE 2^13, 16 queries, 1+8 FRI layers, 54 488 B packed, 3 806 SHA-256 compressions) replicated x1024 per GPU = one PASS.
A STEP = `--passes` (default 256) passes of verify_proof over that 1024-proof batch, issued back to back through the C-ABI
(`config.passes_per_step`): with the driver's `--steps 20` the timed region is about a second instead of 4 ms, so first-use
effects, rank start skew and the one NCCL gather cannot move the number.  Weak scaling: every rank verifies its own
shard; the only exchange is the accept-bitmap gather.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch 1024] [--passes 256] [--mode ref-literal|prover-consistent]

`value`    : whole-job proofs/s with inputs resident in HBM (CUDA events on the launching stream, max over ranks).
`e2e`      : the same metric through the C-ABI with HOST (pinned) buffers: H2D of every pass's batch (compact transport form) + D2H of
             its bitmap inside the timed region; `e2e_packed`: the same on fixed-stride packed records; `e2e_wit`: from `.wit` text.
`roofline` : the dominant kernel against the resource that binds it, the INT32 ALU pipe: ALU-pipe warp instructions per launch (ncu,
             profiles/step_pipe_counts.json) / CUDA-event launch duration / measured ALU issue rate; HBM figures beside it.
`cpu_baseline` : the C oracle (a port of the .simf programs; the reference binary cannot be built here) on the host cores;
             `cpu_baseline_fast`: the same port with SHA-NI / word-wise absorbs / Mersenne folding.
`configs`  : bounded runs of BASELINE configs 1, 3, 4, 5 (bench_sub.py), each with its own parity check.
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALG_BYTES_PER_PROOF = 54488
COMPRESSIONS_PER_PROOF = 3806      # SURVEY.md section 8d (46 channel + 880 trace/CP decommit + 2880 FRI)
MERKLE_COMPRESSIONS_PER_PROOF = 3760
LITERAL_OPS_PER_COMPRESSION = 2296  # FIPS 180-4 literal: 64*26 + 48*13 + 8
LITERAL_OPS_PER_PROOF = 3806 * 2296 + 65486 * 6 + 55220 * 4
METRIC = "stwo_proofs_verified_per_s"


def workload_name(n):
    """One string for both arms (the driver compares `config.workload` of the two lines)."""
    return f"stwo-verifier proof.json witness (prod preset: LDE 2^13, 16 queries, 1+8 FRI layers) replicated x{n} per GPU"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def load_workload(S, batch, mode_name):
    import numpy as np

    mode = S.MODE_REF_LITERAL if mode_name == "ref-literal" else S.MODE_PROVER_CONSISTENT
    cfg = S.stwo_config("prod", mode)
    text = open(os.path.join(ROOT, "tests", "golden", "stwo_proof_prod.wit")).read()
    packed, bad = S.witness.pack_stwo_wits([text], cfg)
    assert not bad[0]
    lo = S.stwo_layout(cfg)
    assert lo.algorithmic_bytes == ALG_BYTES_PER_PROOF
    return cfg, lo, np.tile(packed, batch)


def csrc_sha16():
    d = os.path.join(ROOT, "stark-symphony_b200", "csrc")
    h = hashlib.sha256()
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh")):
            h.update(f.encode())
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def load_pipe_counts(mode_name):
    """ncu-measured warp-instruction counts per kernel launch (profiles/step_pipe_counts.json, written by profiles/pipe_counts.py)."""
    try:
        doc = json.load(open(os.path.join(ROOT, "profiles", "step_pipe_counts.json")))
        m = doc["modes"][mode_name]
        return {"kernels": m["kernels"], "proofs_per_launch": m["proofs_per_launch"], "source": m["source"],
                "matches_build": doc.get("csrc_sha16") == csrc_sha16(), "csrc_sha16": doc.get("csrc_sha16")}
    except Exception as e:  # pragma: no cover
        log(f"profiles/step_pipe_counts.json unavailable: {e}")
        return None


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on all host cores
# ---------------------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(torch, local_rank):
    """Multi-GPU runs feed every GPU from pinned host memory at link rate: pin this rank to the CPUs of its GPU's NUMA node before any
    buffer is allocated (first-touch then places the pinned pages next to the GPU's PCIe root).  Returns the node or None."""
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        bus = f"{getattr(pr, 'pci_domain_id', 0):04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo_, _, hi_ = part.partition("-")
            cpus.update(range(int(lo_), int(hi_ or lo_) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def cpu_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def oracle_run(orc, ocfg, packed, n, threads, proofs_total):
    """Verify `proofs_total` proofs (cycling through the n-proof batch) on `threads` threads; returns (seconds, proofs done)."""
    from oracle import oracle as O

    per = (proofs_total + threads - 1) // threads
    ptr = packed.ctypes.data_as(O.u32p)

    def work(t):
        done, begin = 0, (t * per) % n
        while done < per:
            m = min(per - done, n - begin)
            orc.lib.oracle_stwo_verify_batch(C.byref(ocfg), ptr, begin, begin + m, None, None, None)
            done += m
            begin = (begin + m) % n

    ts = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return time.perf_counter() - t0, per * threads


def cpu_baseline(cfg, packed, n, budget_s=3.0, fast=False):
    from oracle import oracle as O

    orc = O.Oracle()
    ocfg = O.StwoConfig(cfg.trace_log, cfg.lde_log, cfg.n_queries, cfg.n_fri_layers, cfg.mode, cfg.n_columns, cfg.pow_target)
    threads = cpu_threads()
    level = orc.set_fast(fast)
    try:
        t1, _ = oracle_run(orc, ocfg, packed, n, 1, 8)
        per_proof = t1 / 8
        total = int(max(threads * 8, min(budget_s * threads / per_proof, 256 * n)))
        dt, done = oracle_run(orc, ocfg, packed, n, threads, total)
    finally:
        orc.set_fast(False)
    how = ("oracle/ssym_oracle.c -O3, the literal port: byte-at-a-time SHA-256 absorbs, scalar compression, `%` reductions" if not fast else
           "oracle/ssym_oracle.c -O3 in fast mode (oracle_set_fast_sha): 4-byte absorbs, block-wise padding, Mersenne folding, "
           + ("SHA-NI compressions" if level == 2 else "scalar compressions (this host has no SHA-NI)") + "; bit-identical results "
           "(tests/test_oracle_fixtures.py::test_fast_sha_mode_is_bit_identical)")
    return {"value": done / dt, "unit": "proofs/s", "cores": threads, "kind": "port",
            "sample": f"{done} proofs (cycling the same {n}-proof batch) on {threads} threads, {dt:.2f} s wall; {how} "
                      "(the reference's `simfony run` cannot be built here: no Rust, un-vendored crates)",
            "compressions_per_s": done * COMPRESSIONS_PER_PROOF / dt}


def run_reference(args):
    """--impl reference: the reference's CPU path = the oracle port (see cpu_baseline.kind), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np

    from oracle import oracle as O
    from oracle import witparse as W

    mode = O.MODE_REF_LITERAL if args.mode == "ref-literal" else O.MODE_PROVER_CONSISTENT
    ocfg = O.make_config("prod", mode)
    wit = W.load_wit(open(os.path.join(ROOT, "tests", "golden", "stwo_proof_prod.wit")).read())
    one, _ = W.pack_stwo(wit, 16, 8, 13)
    n = args.batch
    packed = np.tile(one, n)
    orc = O.Oracle()
    threads = cpu_threads()
    t1, _ = oracle_run(orc, ocfg, packed, n, 1, 8)
    per_proof = t1 / 8
    budget = 60.0
    m = int(max(threads, min(n * args.passes, budget * threads / ((args.steps + args.warmup) * per_proof))))
    for _ in range(args.warmup):
        oracle_run(orc, ocfg, packed, n, threads, m)
    t0 = time.perf_counter()
    done = 0
    for _ in range(args.steps):
        _, d = oracle_run(orc, ocfg, packed, n, threads, m)
        done += d
    dt = time.perf_counter() - t0
    value = done / dt
    sample = (f"{done // args.steps} proofs per step (a bounded sample of the step's {args.passes} x {n} proofs) x {args.steps} steps on {threads} threads; "
              "oracle/ssym_oracle.c, the literal port (byte-at-a-time SHA-256)")
    # the same sample with the fast CPU path, for context (not the arm's value)
    fast = None
    try:
        level = orc.set_fast(True)
        dtf, df = oracle_run(orc, ocfg, packed, n, threads, m * 4)
        fast = {"value": df / dtf, "unit": "proofs/s", "cores": threads, "kind": "port", "sha_ni": level == 2,
                "sample": f"{df} proofs on {threads} threads, fast mode of the same port (SHA-NI / word-wise absorbs / Mersenne folding)"}
    finally:
        orc.set_fast(False)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "proofs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
        "data": "reference fixture stwo-verifier/tests/data/proof.json (via generate_wit.py) replicated (BASELINE configs[1])",
        "config": {"workload": workload_name(n), "mode": args.mode, "batch_per_gpu": n, "passes_per_step": args.passes,
                   "sampled_proofs_per_step": done // args.steps},
        "cpu_baseline": {"value": value, "unit": "proofs/s", "cores": threads, "kind": "port", "sample": sample},
        "cpu_baseline_fast": fast,
        "e2e": {"value": value, "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ---------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """NVML SM clock / throttle reasons / power, sampled every `interval` s on a side thread (>= 0.2 s: eight ranks polling NVML at
    50 Hz showed up in the round-1 scaling numbers)."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake"}

    def __init__(self, torch_device_index, interval=0.2):
        self.samples, self.reasons, self.power = [], set(), []
        self.interval = interval
        self.stop = threading.Event()
        self.ok = False
        self.max_mhz = None
        try:
            import pynvml
            import torch

            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(torch_device_index).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            try:
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, "encode") else uuid)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(torch_device_index)
            self.nv = pynvml
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            log(f"clock sampling unavailable: {e}")
        self.thread = threading.Thread(target=self.run, daemon=True)

    def sample(self):
        nv = self.nv
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            try:
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for bit, name in self.REASONS.items():
                if r & bit:
                    self.reasons.add(name)
            self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
        except Exception:
            pass

    def run(self):
        while not self.stop.wait(self.interval):
            self.sample()

    def __enter__(self):
        if self.ok:
            self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        if self.ok:
            self.thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s),
                "power_w_max": max(self.power) if self.power else None}


# ---------------------------------------------------------------------------------------------------------------
def emit(line):
    """The ONE JSON line of the contract, on the process's real stdout (see main(): fd 1 is parked on stderr meanwhile)."""
    out = os.fdopen(os.dup(_REAL_STDOUT), "w") if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


_REAL_STDOUT = None


def max_over_ranks(torch, dist, world, values):
    t = torch.tensor(list(values), dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.cpu()]


def gather_ranks(torch, dist, world, values):
    """-> list over ranks of the per-rank value lists."""
    t = torch.tensor(list(values), dtype=torch.float64, device="cuda")
    if world == 1:
        return [[float(x) for x in t.cpu()]]
    out = torch.zeros((world, t.numel()), dtype=torch.float64, device="cuda")
    dist.all_gather_into_tensor(out.view(-1), t)
    return [[float(x) for x in row] for row in out.cpu()]


def main():
    # Libraries (NCCL's "NCCL version ..." banner, for one) write to fd 1: keep stdout clean for the single JSON line.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="proofs per GPU per pass (BASELINE config: 1024)")
    ap.add_argument("--passes", type=int, default=256, help="passes over the batch per step (one C-ABI call each); sizes the timed region")
    ap.add_argument("--e2e-passes", type=int, default=64, help="passes per step of the host-buffer (e2e) legs")
    ap.add_argument("--mode", default="ref-literal", choices=["ref-literal", "prover-consistent"])
    ap.add_argument("--copies", type=int, default=8, help="distinct device copies of the batch rotated through (defeats L2 reuse)")
    ap.add_argument("--pipeline", type=int, default=8, help="batches in flight per GPU (ssym_set_pipeline_depth); 1 = strictly serial passes")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the bounded runs of BASELINE configs 1, 3, 4, 5")
    ap.add_argument("--headline-only", action="store_true", help="device-resident headline loop only (profiling runs)")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)
    args.passes = max(1, args.passes)
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        log(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libssym has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa_node = bind_to_gpu_numa_node(torch, local_rank) if world > 1 and not os.environ.get("SSYM_NO_NUMA_BIND") else None
    if world > 1:
        log(f"rank {rank}: GPU {local_rank} NUMA node {numa_node}, {len(os.sched_getaffinity(0))} CPUs")
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))

    import stark_symphony_b200 as S

    n, R = args.batch, args.passes
    cfg, lo, host_batch = load_workload(S, n, args.mode)
    ver = S.Verifier(local_rank)
    stream = torch.cuda.Stream()  # a real (non-default) stream: the library and the timing events share it
    torch.cuda.set_stream(stream)
    ver.set_stream(stream.cuda_stream)

    # `copies` distinct device copies of the batch, rotated: copies * n * 54.5 KB > 126 MB L2, so no pass finds its input in L2
    depth = max(1, min(8, args.pipeline))
    copies = max(1, args.copies, depth)
    dev = [torch.from_numpy(host_batch.view(np.int32)).cuda() for _ in range(copies)]
    ver.set_pipeline_depth(depth)
    # outputs: one bitmap row per pass of the timed region (never reused inside it), one status buffer per in-flight batch
    words = (n + 31) // 32
    rows = max(args.steps, args.warmup) * R
    accept_all = torch.zeros((rows, words), dtype=torch.int32, device="cuda")
    gathered = torch.zeros((world, rows, words), dtype=torch.int32, device="cuda") if world > 1 else None
    statuses = [torch.zeros(n, dtype=torch.int32, device="cuda") for _ in range(depth)]
    total_n = n * world

    def step(k, c=cfg, acc=accept_all, sts=statuses):
        """One step = R passes of verify_proof over the 1024-proof batch (asynchronous; up to `depth` passes in flight)."""
        for p in range(k * R, (k + 1) * R):
            ver.stwo_verify_batch(dev[p % copies], c, n, accept_out=acc[p % acc.shape[0]], status_out=sts[p % depth])

    def finish(ev_join=None):
        """Order all in-flight batches into the timing stream, then the job's only exchange: the accept-bitmap gather."""
        ver.join()
        if ev_join is not None:
            ev_join.record(stream)
        if world > 1:
            dist.all_gather_into_tensor(gathered.view(-1), accept_all.view(-1))

    int32_lanes, probe_ms = ver.int32_peak_probe()
    for k in range(args.warmup):  # >= 3 steps = >= 3 R passes >> 2 x depth: every lane's scratch exists, NCCL is connected
        step(k)
    finish()
    torch.cuda.synchronize()
    launches0 = ver.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    with ClockSampler(local_rank) as clocks:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        clocks.sample() if clocks.ok else None
        h0 = time.perf_counter()
        ev[0].record(stream)
        for k in range(args.steps):
            step(k)
        h1 = time.perf_counter()
        finish(ev[1])
        ev[2].record(stream)
        torch.cuda.synchronize()
        h2 = time.perf_counter()
        if world > 1:
            dist.barrier()
    launches = ver.launch_count - launches0
    ms = ev[0].elapsed_time(ev[2])
    per_rank = gather_ranks(torch, dist, world, [ms, ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), (h1 - h0) * 1e3, (h2 - h0) * 1e3])
    ms_max = max(r[0] for r in per_rank)
    value = total_n * R * args.steps / (ms_max * 1e-3)
    pass_ms = ms_max / (args.steps * R)

    # correctness of what was timed (outside the timed region): every pass's bitmap, every rank's statuses against the oracle
    st = statuses[(args.steps * R - 1) % depth].cpu().numpy().view(np.uint32)
    used_rows = accept_all[: args.steps * R].cpu().numpy()
    assert (used_rows == used_rows[0]).all(), "passes disagree"
    last_row = args.steps * R - 1
    last = (gathered[:, last_row, :] if world > 1 else accept_all[last_row]).contiguous().cpu().numpy()
    bits = np.concatenate([np.unpackbits(r.view(np.uint8), bitorder="little")[:n] for r in last.reshape(world, words)])
    accepted = int(bits.sum())
    if rank == 0:
        from oracle import oracle as O

        orc = O.Oracle()
        ocfg = O.StwoConfig(cfg.trace_log, cfg.lde_log, cfg.n_queries, cfg.n_fri_layers, cfg.mode, cfg.n_columns, cfg.pow_target)
        _, o_status, _ = orc.stwo_verify_batch(ocfg, host_batch, 4)
        assert (st[:4] == o_status).all() and (st == st[0]).all(), "GPU status differs from the oracle"
        assert accepted == (total_n if args.mode == "prover-consistent" else 0)

    line = {
        "metric": METRIC, "value": value, "unit": "proofs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
        "data": "reference fixture stwo-verifier/tests/data/proof.json (via generate_wit.py) replicated (BASELINE configs[1]); distinct GPU-proven proofs: configs.c3 / c5",
        "config": {"workload": workload_name(n), "mode": args.mode, "batch_per_gpu": n, "passes_per_step": R, "proofs_per_step": total_n * R,
                   "ms_per_pass": pass_ms, "accepted": accepted,
                   "step": f"one step = {R} passes of verify_proof over the {n}-proof batch, one ssym_stwo_verify_batch call per pass (timed region = steps x passes calls)",
                   "l2": f"rotating {copies} distinct device copies of the batch ({copies * n * lo.stride_words * 4 / 1e6:.0f} MB > 126 MB L2)",
                   "pipeline": f"{depth} batches in flight per GPU (ssym_set_pipeline_depth); per-kernel times below are per launch, kernels of consecutive passes overlap" if depth > 1 else "serial passes",
                   "parallelism": f"proof-sharded x{world} (one process per GPU, no data-path collective), one NCCL all_gather of the accept bitmaps of all passes inside the timed region" if world > 1 else "single GPU",
                   "numa": (f"every rank pinned to the CPUs of its GPU's NUMA node (rank 0: node {numa_node})" if numa_node is not None else "no NUMA binding")},
        "merkle_hashes_per_s": value * MERKLE_COMPRESSIONS_PER_PROOF / 2.0,
        "sha256_compressions_per_s": value * COMPRESSIONS_PER_PROOF,
        "gpu_launches": int(launches),
        "timing_per_rank": {"columns": ["timed_ms (events: first launch -> after gather)", "compute_ms (first launch -> all batches joined)",
                                        "allgather_ms (device)", "launch_loop_ms (host: issuing all calls)", "host_wall_ms (launch -> synchronized)"],
                            "ranks": per_rank,
                            "note": "the headline uses max over ranks of column 0; a launch loop as long as the timed region means the host, not the GPU, sets the rate"},
        "clocks": clocks.summary(),
    }
    if args.headline_only:
        if rank == 0:
            emit(line)
        ver.close()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- the same pipelined loop under the other semantics ---------------------------------------------
    # PROVER_CONSISTENT accepts the fixture, and accepted proofs are where the shared-node Merkle schedule applies.
    other_mode = S.MODE_PROVER_CONSISTENT if args.mode == "ref-literal" else S.MODE_REF_LITERAL
    other_name = "prover-consistent" if args.mode == "ref-literal" else "ref-literal"
    cfg_other = S.stwo_config("prod", other_mode)
    other_steps = max(3, args.steps // 4)
    accept_other = torch.zeros((R, words), dtype=torch.int32, device="cuda")
    statuses_other = [torch.zeros(n, dtype=torch.int32, device="cuda") for _ in range(depth)]
    for k in range(2):
        step(k, cfg_other, accept_other, statuses_other)
    ver.join()
    torch.cuda.synchronize()
    o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    o0.record(stream)
    oh0 = time.perf_counter()
    for k in range(other_steps):
        step(k, cfg_other, accept_other, statuses_other)
    other_launch_ms = (time.perf_counter() - oh0) * 1e3
    ver.join()
    o1.record(stream)
    torch.cuda.synchronize()
    other_ms = max_over_ranks(torch, dist, world, [o0.elapsed_time(o1)])[0]
    other_value = total_n * R * other_steps / (other_ms * 1e-3)
    other_accepted = int(np.unpackbits(accept_other[R - 1].cpu().numpy().view(np.uint8), bitorder="little")[:n].sum())
    assert other_accepted == (n if other_mode == S.MODE_PROVER_CONSISTENT else 0)

    # ---- per-kernel durations for the roofline: the same passes issued strictly serially (depth 1), CUDA events around every kernel ----
    ver.set_pipeline_depth(1)
    prof_by_mode = {}
    serial_ms = {}
    for name, c in ((args.mode, cfg), (other_name, cfg_other)):
        ver.profile_read()
        ver.profile_enable(True)
        serial_passes = 200
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record(stream)
        for p in range(serial_passes):
            ver.stwo_verify_batch(dev[p % copies], c, n, accept_out=accept_other[p % R], status_out=statuses_other[0])
        e3.record(stream)
        torch.cuda.synchronize()
        serial_ms[name] = e2.elapsed_time(e3) / serial_passes
        ver.profile_enable(False)
        prof_by_mode[name] = ver.profile_read()
    ver.set_pipeline_depth(depth)
    prof = prof_by_mode[args.mode]

    # ---- end to end: host (pinned) buffers through the same C-ABI calls, copies inside the timed region ---------------------------
    ver.set_stream(None)
    Re = max(1, args.e2e_passes)
    e2e_steps = args.steps
    acc_host = torch.empty(words, dtype=torch.int32).pin_memory().numpy().view(np.uint32)
    acc_rows = torch.zeros((Re, words), dtype=torch.int32).pin_memory()
    acc_rows_np = acc_rows.numpy().view(np.uint32)
    pinned = torch.empty(host_batch.size, dtype=torch.int32).pin_memory()
    pinned.numpy()[:] = host_batch.view(np.int32)
    host_view = pinned.numpy().view(np.uint32)

    # compact transport form (include/ssym.h), version 3: produced OUTSIDE the timed region by ssym_stwo_compact_pack_gpu (the GPU finds the
    # siblings that are nodes of other queries' paths, the host assembles the records); its cost is reported.  Records with derived siblings are
    # bound to the semantics they were packed under, so each mode gets its own blob.
    bound_words = int(S.load().ssym_stwo_compact_bound(C.byref(cfg), n))

    def pack_compact(c):
        pinned_buf = torch.empty(bound_words, dtype=torch.int32).pin_memory()
        S.witness.compact_stwo(host_batch, c, out=pinned_buf.numpy().view(np.uint32), ver=ver)  # warm (page faults of the output)
        t0 = time.perf_counter()
        full, off_np = S.witness.compact_stwo(host_batch, c, out=pinned_buf.numpy().view(np.uint32), ver=ver)
        secs = time.perf_counter() - t0
        words = int(off_np[n])
        off_t = torch.empty(n + 1, dtype=torch.int64).pin_memory()
        off_t.numpy()[:] = off_np.view(np.int64)
        return {"pinned": pinned_buf, "blob": full[:words], "words": words, "off": off_t, "off_view": off_t.numpy().view(np.uint64), "pack_s": secs,
                "derived_per_proof": int(full[4]) if int(full[2]) == 0x33435353 else 0}

    # The headline leg ships the records packed under prover-consistent — the smallest lossless form of the batch (30.7 KB per proof: every tree's
    # siblings are derivable where paths meet) — whatever mode it verifies under.  Under ref-literal the library then takes the cross path
    # (include/ssym.h; launch_stwo_verify_cross): one transcript, the records' own FRI evaluations and chains to complete the packed records, then the
    # verification proper; bit-identical statuses (tests/test_compact.py).  `same_mode_records` is the leg on records packed under the call's own mode.
    cp = pack_compact(cfg)
    cp_other = pack_compact(cfg_other)
    t0 = time.perf_counter()
    S.witness.compact_stwo(host_batch, cfg)  # version 2: host only (no hashing), for comparison
    pack_v2_s = time.perf_counter() - t0
    c_pinned, c_blob, c_words, c_off, c_off_view, pack_s = cp["pinned"], cp["blob"], cp["words"], cp["off"], cp["off_view"], cp["pack_s"]

    def host_leg(call, nbytes):
        """-> (async proofs/s, sync proofs/s, launches per pass, plain-copy GB/s alone, plain-copy GB/s with all ranks copying) for one host-buffer entry point."""
        for _ in range(3):
            call(acc_host)
        ref_bits = acc_host.copy()
        l0 = ver.launch_count
        sync_passes = max(8, Re // 4)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(sync_passes):
            call(acc_host)  # synchronous: returns with the bitmap in host memory
        sync_s = time.perf_counter() - t0
        per_pass = (ver.launch_count - l0) // sync_passes
        # throughput mode of the same call (ssym_set_host_async): each call enqueues its H2D + kernels + D2H and returns; one synchronize per step;
        # every pass of a step has its own pinned bitmap row, checked after the timed region
        ver.set_host_async(True)
        for k in range(4):
            call(acc_rows_np[k % Re])
        ver.synchronize()
        acc_rows_np[:] = 0xA5A5A5A5
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            for p in range(Re):
                call(acc_rows_np[p])
            ver.synchronize()
        async_s = time.perf_counter() - t0
        ver.set_host_async(False)
        assert (acc_rows_np == ref_bits[None, :]).all(), "asynchronous host-buffer results differ from the synchronous call"
        a, s_ = max_over_ranks(torch, dist, world, [async_s, sync_s])
        return total_n * Re * e2e_steps / a, total_n * sync_passes / s_, int(per_pass), ref_bits

    def plain_copy_gbs(src_tensor, nbytes, concurrent):
        """Plain pinned H2D copies of the same bytes: the link ceiling of a host leg.  concurrent: all ranks copy at the same time (barrier)."""
        dst = torch.empty(src_tensor.numel(), dtype=src_tensor.dtype, device="cuda")
        reps = max(4, int(0.25 / (nbytes / 50e9)))
        dst.copy_(src_tensor, non_blocking=True)
        torch.cuda.synchronize()
        if concurrent and world > 1:
            dist.barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(reps):
            dst.copy_(src_tensor, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        return reps * nbytes / (c0.elapsed_time(c1) * 1e-3) / 1e9

    c_bytes = c_words * 4 + c_off.numpy().nbytes
    c_value, c_sync, c_launches, bits_c = host_leg(lambda acc: ver.stwo_verify_compact_batch(c_blob, c_off_view, cfg, accept_out=acc), c_bytes)
    ship = cp if args.mode == "prover-consistent" else cp_other  # the records packed under prover-consistent
    s_bytes = ship["words"] * 4 + ship["off"].numpy().nbytes
    if ship is cp:
        s_value, s_sync, s_launches, bits_s = c_value, c_sync, c_launches, bits_c
    else:
        s_value, s_sync, s_launches, bits_s = host_leg(lambda acc: ver.stwo_verify_compact_batch(ship["blob"], ship["off_view"], cfg, accept_out=acc), s_bytes)
    assert (bits_s == bits_c).all(), "records packed under the other mode verify differently"
    co_bytes = cp_other["words"] * 4 + cp_other["off"].numpy().nbytes
    co_value, co_sync, _, bits_co = host_leg(lambda acc: ver.stwo_verify_compact_batch(cp_other["blob"], cp_other["off_view"], cfg_other, accept_out=acc), co_bytes)
    assert int(np.unpackbits(bits_co.view(np.uint8), bitorder="little")[:n].sum()) == (n if other_mode == S.MODE_PROVER_CONSISTENT else 0)
    p_bytes = n * lo.stride_words * 4
    p_value, p_sync, p_launches, bits_p = host_leg(lambda acc: ver.stwo_verify_batch(host_view, cfg, n, accept_out=acc), p_bytes)
    assert (bits_c == bits_p).all() and (np.unpackbits(bits_p.view(np.uint8), bitorder="little")[:n] == bits[:n]).all(), "host legs disagree with the device leg"
    copy_alone = plain_copy_gbs(ship["pinned"][:ship["words"]], ship["words"] * 4, False)
    if world > 1:
        dist.barrier()
    copy_conc = plain_copy_gbs(ship["pinned"][:ship["words"]], ship["words"] * 4, True)
    conc_ranks = gather_ranks(torch, dist, world, [copy_alone, copy_conc])

    # end to end from the reference's own input format: `.wit` JSON TEXT in pinned host memory -> accept bits
    # (ssym_stwo_verify_wit_batch: text H2D, GPU tokeniser + packer, verifier, D2H bitmap; 122 KB of text per proof)
    wit_raw = open(os.path.join(ROOT, "tests", "golden", "stwo_proof_prod.wit"), "rb").read()
    n_wit = 8 * n  # 8 192 witnesses = 1 GB of text per call: the synchronous call's exposed first copy / last kernels stay below 10 %
    wit_pinned = torch.empty(len(wit_raw) * n_wit, dtype=torch.uint8).pin_memory()
    wit_np = wit_pinned.numpy()
    wit_np.reshape(n_wit, len(wit_raw))[:] = np.frombuffer(wit_raw, dtype=np.uint8)
    wit_offsets = (np.arange(n_wit + 1, dtype=np.uint64) * np.uint64(len(wit_raw)))
    acc_wit = torch.empty((n_wit + 31) // 32, dtype=torch.int32).pin_memory().numpy().view(np.uint32)
    wit_calls = 8
    ver.stwo_verify_wit_batch(wit_np, wit_offsets, cfg, accept_out=acc_wit)
    wl0 = ver.launch_count
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(wit_calls):
        ver.stwo_verify_wit_batch(wit_np, wit_offsets, cfg, accept_out=acc_wit)
    wit_s = max_over_ranks(torch, dist, world, [time.perf_counter() - t0])[0]
    wit_launches = ver.launch_count - wl0
    assert (np.unpackbits(acc_wit.view(np.uint8), bitorder="little")[:n_wit] == bits[0]).all(), "text path disagrees with the packed path"
    wit_value = n_wit * world * wit_calls / wit_s
    del wit_pinned, wit_np

    # ---- bounded runs of BASELINE configs 1, 3, 4 (single GPU only) and 5 (every N) -------------------------------------------------
    sub = {}
    if not args.no_configs:
        import bench_sub

        ver.set_stream(stream.cuda_stream)
        ver.set_pipeline_depth(1)
        sub = bench_sub.run_all(S, ver, stream, rank, world, int32_lanes, single_gpu_configs=(world == 1))
        ver.set_pipeline_depth(depth)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json (driver-measured copy bandwidth)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
        mk_name = "stwo_merkle"
        mk_ms, mk_n = prof.get(mk_name, (0.0, 0))
        mk_avg_ms = mk_ms / max(mk_n, 1)
        kernel_ms = {k: v[0] / max(v[1], 1) for k, v in prof.items()}
        share = mk_ms / max(sum(v[0] for v in prof.values()), 1e-9)
        hbm_achieved = n * ALG_BYTES_PER_PROOF / (mk_avg_ms * 1e-3) / 1e9 if mk_avg_ms else 0.0
        lit_ops = n * MERKLE_COMPRESSIONS_PER_PROOF * LITERAL_OPS_PER_COMPRESSION / (mk_avg_ms * 1e-3) if mk_avg_ms else 0.0
        warp_peak = int32_lanes / 32.0  # ALU-pipe warp instructions per second, measured (ssym_int32_peak_probe)
        counts = load_pipe_counts(args.mode)
        roof = {"bound": "int32_alu", "kernel": "stwo_merkle_kernel", "unit": "G warp-instructions/s (ALU pipe)", "peak": warp_peak / 1e9,
                "peak_source": f"ssym_int32_peak_probe, measured in this run: {int32_lanes / 1e12:.2f} T SHF/LOP3/IADD3 lanes/s = 148 SMs x 64 lanes x clock ({probe_ms:.2f} ms probe)",
                "avg_launch_ms": mk_avg_ms, "kernel_share_of_step": share}
        if counts and "stwo_merkle_kernel" in counts["kernels"] and mk_avg_ms:
            scale = n / counts["proofs_per_launch"]
            k3 = counts["kernels"]["stwo_merkle_kernel"]
            alu = k3["alu_pipe_warp_inst"] * scale
            step_alu = sum(k.get("alu_pipe_warp_inst", 0.0) for k in counts["kernels"].values()) * scale
            roof.update({
                "achieved": alu / (mk_avg_ms * 1e-3) / 1e9, "frac": alu / (mk_avg_ms * 1e-3) / warp_peak,
                "alu_pipe_warp_inst_per_launch": alu, "warp_inst_per_launch": k3.get("warp_inst", 0.0) * scale,
                "fmaheavy_pipe_warp_inst_per_launch": k3.get("fmaheavy_pipe_warp_inst", 0.0) * scale,
                "whole_step_frac": step_alu / (pass_ms * 1e-3) / warp_peak,
                "whole_step_note": "ALU-pipe warp instructions of ALL kernels of one pass / the headline's ms_per_pass (pipelined) / the same peak",
                "serial_step_frac": step_alu / (serial_ms[args.mode] * 1e-3) / warp_peak,
                "traffic": k3.get("dram_read_bytes", 0.0) * scale + k3.get("dram_write_bytes", 0.0) * scale,
                "counts_source": f"profiles/step_pipe_counts.json <- {counts['source']} (ncu, per launch at {counts['proofs_per_launch']} proofs; instruction counts are clock- and profiler-independent)",
                "counts_match_build": counts["matches_build"],
            })
        else:
            roof.update({"achieved": None, "frac": None, "traffic": None, "counts_source": "profiles/step_pipe_counts.json missing"})
        other_counts = load_pipe_counts(other_name)
        other_frac = None
        if other_counts:
            o_alu = sum(k.get("alu_pipe_warp_inst", 0.0) for k in other_counts["kernels"].values()) * n / other_counts["proofs_per_launch"]
            other_frac = o_alu / (other_ms / (other_steps * R) * 1e-3) / warp_peak
        roof["hbm"] = {"achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak if hbm_peak else None,
                       "algorithmic_bytes_per_launch": n * ALG_BYTES_PER_PROOF, "peak_source": peak_src,
                       "note": "not the binding resource: 170 integer ops per byte"}
        roof["literal_ops"] = {"achieved": lit_ops / 1e12, "peak": int32_lanes / 1e12, "unit": "Tops/s", "frac": lit_ops / int32_lanes if int32_lanes else None,
                               "convention": "SURVEY 8d: FIPS-180-4-literal 2296 ops x compressions / kernel time over measured instruction lanes/s; LOP3/IADD3 fusion and "
                                             "adds issued on the FMA pipe make > 1.0 possible — `frac` above is the instruction-level figure"}
        line.update({
            "other_mode": {"mode": other_name, "value": other_value, "unit": "proofs/s", "steps": other_steps, "accepted_per_gpu": other_accepted,
                           "kernel_ms": {k: v[0] / max(v[1], 1) for k, v in prof_by_mode[other_name].items()}, "serial_ms_per_pass": serial_ms[other_name],
                           "whole_step_frac_int32_alu": other_frac, "timed_ms": other_ms, "launch_loop_ms": other_launch_ms,
                           "note": "same pipelined loop, same batch, the other semantics switch; under prover-consistent the fixture is ACCEPTED and the Merkle paths "
                                   "of a tree share the nodes above the height where they meet (hashed once, results per query identical: DESIGN.md section 4), "
                                   "so fewer compressions are executed than the reference's per-query count"},
            "e2e": {"value": s_value, "unit": "proofs/s", "h2d_bytes_per_step": int(s_bytes * Re), "d2h_bytes_per_step": int(acc_host.nbytes * Re),
                    "steps": e2e_steps, "passes_per_step": Re, "h2d_gbs_achieved": s_value / world * s_bytes / n / 1e9,
                    "h2d_gbs_plain_copy": copy_alone, "h2d_gbs_plain_copy_concurrent": min(r[1] for r in conc_ranks),
                    "h2d_gbs_plain_copy_concurrent_per_rank": [r[1] for r in conc_ranks], "h2d_gbs_plain_copy_concurrent_aggregate": sum(r[1] for r in conc_ranks),
                    "frac_of_concurrent_copy": (s_value / world * s_bytes / n / 1e9) / max(min(r[1] for r in conc_ranks), 1e-9),
                    "sync_call_value": s_sync, "bytes_per_proof": ship["words"] * 4 / n, "packed_bytes_per_proof": lo.stride_words * 4,
                    "gpu_launches_per_pass": s_launches,
                    "derived_siblings_per_proof": ship["derived_per_proof"], "records_packed_under": "prover-consistent", "verified_under": args.mode,
                    "same_mode_records": {"value": c_value, "sync_call_value": c_sync, "bytes_per_proof": c_words * 4 / n, "records_packed_under": args.mode,
                                          "derived_siblings_per_proof": cp["derived_per_proof"], "gpu_launches_per_pass": c_launches,
                                          "note": "the same leg on records packed under the call's own mode (one pass over the kernels)"},
                    "host_pack": {"proofs_per_s": n / pack_s, "seconds_per_1024": pack_s, "inside_timed_region": False,
                                  "version2_host_only_proofs_per_s_per_core": n / pack_v2_s,
                                  "note": "ssym_stwo_compact_pack_gpu turns packed records into version 3 compact records BEFORE the clock starts: H2D of the packed "
                                          "records, the verifier's kernels in scan mode (which siblings are nodes of other paths), D2H of one byte per sibling slot, "
                                          "assembly on one host thread.  A producer that emits compact records directly (a prover: it holds the trees) pays nothing; "
                                          "version 2 records (no hashing, host only: ssym_stwo_compact_pack) cost `version2_host_only_proofs_per_s_per_core`"},
                    "other_mode": {"mode": other_name, "value": co_value, "sync_call_value": co_sync, "bytes_per_proof": cp_other["words"] * 4 / n,
                                   "derived_siblings_per_proof": cp_other["derived_per_proof"], "h2d_gbs_achieved": co_value / world * co_bytes / n / 1e9,
                                   "note": "the same leg under the other semantics, on records packed under it.  Which siblings can be left out depends on the proof "
                                           "being consistent with the verifier: under ref-literal the fixture's FRI evaluations are not the ones its FRI trees were built "
                                           "from (finding F1, DESIGN.md section 1), so only the trace and composition trees have derivable siblings; under "
                                           "prover-consistent all eleven trees do and the record is the size of upstream stwo's minimal decommitment"},
                    "e2e_packed_value": p_value,
                    "note": "ssym_stwo_verify_compact_batch(SSYM_MEM_HOST) on pinned host buffers holding the batch in the compact transport form "
                            "(include/ssym.h, version 3: per Merkle tree every distinct 32-byte sibling once, none at all where another query's path computes it; "
                            "lossless for any record; expanded on the GPU by stwo_expand_kernel + the Merkle kernel itself).  The records are the ones packed under "
                            "prover-consistent (the size of upstream stwo's minimal decommitment); verified under ref-literal they take one transcript, the records' own "
                            "FRI chains to complete them and then the verification proper (launch_stwo_verify_cross), statuses bit-identical to the packed path: "
                            "chunked multi-buffered H2D -> expand -> verifier kernels -> D2H bitmap, every "
                            "pass's copies inside the timed region.  PACKING IS OUTSIDE THE CLOCK (host_pack).  `value`: calls enqueued back to back "
                            "(ssym_set_host_async), one synchronize per step; `sync_call_value`: each call returns with its bitmap in host memory.  Bound by the host link; "
                            "`e2e_packed` is the same measurement on the reference-shaped fixed-stride packed records (no host packing at all)"},
            "e2e_packed": {"value": p_value, "unit": "proofs/s", "h2d_bytes_per_step": int(p_bytes * Re), "d2h_bytes_per_step": int(acc_host.nbytes * Re),
                           "steps": e2e_steps, "passes_per_step": Re, "h2d_gbs_achieved": p_value / world * lo.stride_words * 4 / 1e9, "h2d_gbs_plain_copy": copy_alone,
                           "sync_call_value": p_sync, "gpu_launches_per_pass": p_launches,
                           "note": "ssym_stwo_verify_batch(SSYM_MEM_HOST) on pinned host buffers holding fixed-stride packed records (the witness's values one to one): chunked "
                                   "multi-buffered H2D -> kernels -> D2H bitmap inside the timed region"},
            "e2e_wit": {"value": wit_value, "unit": "proofs/s", "h2d_bytes_per_step": int(len(wit_raw) * n_wit), "d2h_bytes_per_step": int(acc_wit.nbytes),
                        "steps": wit_calls, "proofs_per_step": n_wit, "h2d_gbs_achieved": wit_value / world * len(wit_raw) / 1e9, "gpu_launches": int(wit_launches),
                        "note": "ssym_stwo_verify_wit_batch(SSYM_MEM_HOST): the reference's own input, `.wit` JSON text (122 KB per proof, the file `simfony run "
                                "--witness` reads) in pinned host memory -> GPU tokeniser/packer -> verifier -> bitmap; synchronous calls, text H2D of chunk k+1 under "
                                "the kernels of chunk k"},
            "kernel_ms": kernel_ms, "serial_ms_per_pass": serial_ms[args.mode],
            "kernel_ms_note": "per-launch CUDA-event durations from a strictly serial run (pipeline depth 1) of 200 passes; the headline "
                              "`value` keeps `pipeline` batches in flight so kernels of consecutive passes overlap",
            "roofline": roof,
            "configs": sub,
        })
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(cfg, host_batch, n)
            line["cpu_baseline_fast"] = cpu_baseline(cfg, host_batch, n, budget_s=2.0, fast=True)
        emit(line)
    ver.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
