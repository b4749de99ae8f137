#!/usr/bin/env python3
"""ncu per-launch metrics CSV -> profiles/step_pipe_counts.json, the file bench.py reads its instruction-level roofline numerators from.

  ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fmaheavy.sum,dram__bytes_read.sum,dram__bytes_write.sum \
      --clock-control none -k regex:stwo_ -s <warm-up launches> -c 12 --csv --log-file gpurun_out/r02_step_metrics.csv \
      python bench.py --steps 1 --warmup 1 --passes 4 --pipeline 1 --no-cpu-baseline --no-configs --headline-only [--mode prover-consistent]
  python profiles/pipe_counts.py profiles/r02_step_metrics.csv [profiles/r02_shared_step_metrics.csv] --proofs 1024[,8192] -o profiles/step_pipe_counts.json

Per kernel (name without template / argument list): the MEDIAN over the captured passes of warp instructions executed, ALU-pipe and FMA-heavy-pipe warp
instructions, DRAM bytes, and the (cold-cache, serialised) duration under ncu.  Instruction counts do not depend on clocks or on the profiler, which
is why they may be measured once per build and divided by the CUDA-event time of a live run; durations under ncu are only good for shares."""
import argparse
import csv
import hashlib
import json
import os
import re
import statistics
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = {
    "smsp__inst_executed.sum": "warp_inst", "sm__inst_executed_pipe_alu.sum": "alu_pipe_warp_inst",
    "sm__inst_executed_pipe_fmaheavy.sum": "fmaheavy_pipe_warp_inst", "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes", "gpu__time_duration.sum": "ncu_duration_ns",
}


def csrc_sha16():
    """Hash of the CUDA sources the counts belong to (bench.py compares it with the tree it runs from)."""
    d = os.path.join(ROOT, "stark-symphony_b200", "csrc")
    h = hashlib.sha256()
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh")):
            h.update(f.encode())
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def kernel_base(name):
    name = re.sub(r"^void\s+", "", name)
    return re.split(r"[<(]", name, 1)[0].strip()


def parse(path):
    rows = [r for r in csv.reader(open(path, newline="")) if len(r) > 10]
    hdr = rows[0]
    per_launch = defaultdict(dict)
    for r in rows[1:]:
        x = dict(zip(hdr, r))
        if x["Metric Name"] in METRICS:
            per_launch[(int(x["ID"]), x["Kernel Name"], x["Grid Size"], x["Block Size"])][METRICS[x["Metric Name"]]] = float(x["Metric Value"].replace(",", ""))
    by_kernel = defaultdict(list)
    for (_, name, grid, block), m in sorted(per_launch.items()):
        by_kernel[kernel_base(name)].append(dict(m, full_name=name, grid=grid, block=block))
    # A pass launches stwo_finalize_kernel exactly once; a kernel launched several times per pass (the two rounds of stwo_merkle_shared_kernel)
    # is reported as its per-pass TOTAL, so that "sum over kernels" is the pass.
    passes = max(1, len(by_kernel.get("stwo_finalize_kernel", [])) or 1)
    out = {}
    for k, launches in by_kernel.items():
        per_pass = max(1, round(len(launches) / passes))
        rec = {"launches_seen": len(launches), "launches_per_pass": per_pass, "full_name": launches[0]["full_name"], "grid": launches[0]["grid"],
               "block": launches[0]["block"]}
        for key in METRICS.values():
            vals = [l[key] for l in launches if key in l]
            if not vals:
                continue
            if per_pass == 1:
                rec[key] = statistics.median(vals)
            else:  # launches of one pass are consecutive: median over passes of the per-pass sum
                sums = [sum(vals[j:j + per_pass]) for j in range(0, len(vals) - per_pass + 1, per_pass)]
                rec[key] = statistics.median(sums)
        out[k] = rec
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv", nargs="+", help="first = REF_LITERAL step (per-query Merkle kernel); optional second = PROVER_CONSISTENT step (shared-node schedule)")
    ap.add_argument("--proofs", default="1024", help="proofs per launch in the captured run(s): one number, or one per csv separated by commas")
    ap.add_argument("-o", "--out", default=os.path.join(ROOT, "profiles", "step_pipe_counts.json"))
    args = ap.parse_args()
    proofs = [int(x) for x in args.proofs.split(",")]
    proofs += [proofs[-1]] * (len(args.csv) - len(proofs))
    doc = {"csrc_sha16": csrc_sha16(),
           "how": "ncu --metrics (see profiles/pipe_counts.py), median per kernel over the captured launches; warp-level instruction counts",
           "modes": {}}
    for name, path, n in zip(("ref-literal", "prover-consistent"), args.csv, proofs):
        doc["modes"][name] = {"proofs_per_launch": n, "source": os.path.relpath(os.path.abspath(path), ROOT), "kernels": parse(path)}
    with open(args.out, "w") as f:
        json.dump(doc, f, indent=1, sort_keys=True)
    for mode, m in doc["modes"].items():
        tot = sum(k.get("alu_pipe_warp_inst", 0) * 1 for k in m["kernels"].values())
        print(mode, {k: int(v.get("alu_pipe_warp_inst", 0)) for k, v in m["kernels"].items()}, "ALU-pipe warp instructions per launch; sum", int(tot))


if __name__ == "__main__":
    main()
