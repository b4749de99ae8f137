#!/usr/bin/env python3
"""profiles/r02_ncu_summary.md from the captures of tools/ncu_r02.sh and a bench line of the same build.
  python profiles/make_r02_summary.py gpurun_out/r02q_bench.json gpurun_out/r02_merkle.ncu-rep gpurun_out/r02_channel.ncu-rep > profiles/r02_ncu_summary.md"""
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

HERE = os.path.dirname(os.path.abspath(__file__))
bench, merkle_rep, channel_rep = sys.argv[1:4]
c = json.load(open(os.path.join(HERE, "step_pipe_counts.json")))
rl, pc = c["modes"]["ref-literal"]["kernels"], c["modes"]["prover-consistent"]["kernels"]
d = json.load(open(bench))
r = d["roofline"]
alu_k3 = rl["stwo_merkle_kernel"]["alu_pipe_warp_inst"]
alu_all = sum(v["alu_pipe_warp_inst"] for v in rl.values())
alu_pc = sum(v["alu_pipe_warp_inst"] for v in pc.values())
ms_k3, peak = d["kernel_ms"]["stwo_merkle"], r["peak"]
ms_pass = d["config"]["ms_per_pass"]
pc_ms = 1024 / d["other_mode"]["value"] * 1e3
dur = {k: v["ncu_duration_ns"] / 1e3 for k, v in rl.items()}
tot = sum(dur.values())
k3 = rl["stwo_merkle_kernel"]

rows = list(csv.reader(subprocess.run(["ncu", "-i", merkle_rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
op = defaultdict(lambda: defaultdict(float))
for row in rows[2:]:
    src = row[ix["Source"]].split()
    if not src:
        continue
    o = (src[0] if not src[0].startswith("@") else src[1]).split(".")[0]
    for k in ("# Samples", "stall_dispatch", "stall_math", "stall_wait", "stall_not_selected", "stall_selected", "stall_long_sb", "stall_no_inst"):
        try:
            op[o][k] += float(row[ix[k]])
        except ValueError:
            pass
S = lambda o, k: f"{int(op[o][k]):,}".replace(",", " ")
total_samples = f"{int(sum(v['# Samples'] for v in op.values())):,}".replace(",", " ")

print(f"""# ncu summary — round 2 build (final: SHA additions with the multiplier in a uniform register, `ADDMODE 8`)

Commands: `tools/ncu_r02.sh` (one GPU, `gpurun`; numbers printed by a run under ncu are not bench values); this file: `profiles/make_r02_summary.py`.  Source
hash of the CUDA tree the counts belong to: `{c['csrc_sha16']}` (`bench.py` compares it with the tree it runs from: `roofline.counts_match_build`).  Files:

| file | what |
|---|---|
| `r02_step_metrics.csv` | every `stwo_*` kernel of three serial REF_LITERAL passes at 1024 proofs: duration, warp instructions, ALU-pipe / FMA-heavy-pipe warp instructions, DRAM bytes |
| `r02_shared_step_metrics.csv` | the same for two PROVER_CONSISTENT passes (shared-node Merkle schedule: plan / round 1 / check / round 2 / resolve) |
| `step_pipe_counts.json` | per-kernel per-pass medians of the two files above (`profiles/pipe_counts.py`) — **`bench.py` and `bench_sub.py` read their roofline numerators from this file** |
| `r02_launches.csv` | launch list (`gpu__time_duration.sum`) of the first 600 launches of `bench.py --steps 2 --warmup 3 --passes 8 --e2e-passes 4` |
| `r02_k3_sass_excerpt.md` | instruction mix and excerpts of the K3 hashing loops from `cuobjdump -sass` |
| `{os.path.basename(bench).replace('.json', '_n1.json') if '_n1' not in bench else os.path.basename(bench)}`, `r02o_reference_arm.json` | the bench line of the same build (`python bench.py --steps 20 --warmup 5`) and the reference arm (`--impl reference`); `r02u_bench_n{{2,4}}.json`, `r02t_bench_n8.json`: the multi-GPU lines; `r02d_bench_n1.json` / `r02o_bench_n1.json`: earlier in the round — before the uniform-register multiplier / before the host-path changes, the draw warp of K1 and the one-warp CTAs of K3 |
| `r02_micro.json`, `r02_config3.json`, `r02_config5_n8.json` | full-size runs of BASELINE configs 4 / 1, 3 and 5 (`bench_micro.py`, `bench_configs.py`) |
| `r02_sanitizer_summary.txt` | `tools/sanitize.sh all`: memcheck, initcheck (87 tests each), racecheck (66), synccheck (31): 0 errors / 0 hazards |

How the bench line's roofline follows from these (recompute with the bench line's `kernel_ms.stwo_merkle` and `roofline.peak`):

* `roofline.frac` = {alu_k3:,.0f} ALU-pipe warp instructions (K3, 1024 proofs) / `kernel_ms.stwo_merkle` / (`int32_peak_probe` lanes/s / 32).  With {ms_k3:.4f} ms and
  {peak * 32 / 1e3:.2f} T lanes/s: {alu_k3 / ms_k3 / 1e6:.1f} / {peak:.1f} G warp-instructions/s = **{alu_k3 / ms_k3 / 1e6 / peak:.3f}** (a lone launch: 5 632 one-warp CTAs on 4 736 slots = 1.19 waves; ncu below).  Before the uniform-register multiplier: 0.733 – 0.755 at 0.241 – 0.248 ms (ncu: 73.2 % / 82.7 %).
* `roofline.whole_step_frac` = {alu_all:,.0f} (all four kernels of a pass) / `config.ms_per_pass` / the same peak = **{alu_all / ms_pass / 1e6 / peak:.3f}** at {ms_pass:.4f} ms per pass (depth-8 pipeline: the
  tails of consecutive launches overlap).
* PROVER_CONSISTENT pass: {alu_pc:,.0f} ALU-pipe warp instructions ({alu_pc / alu_all:.3f} x the per-query schedule: every distinct node hashed once), {pc_ms:.4f} ms per pass pipelined = {alu_pc / pc_ms / 1e6 / peak:.2f}.
* DRAM traffic of K3: {k3['dram_read_bytes'] / 1e6:.2f} MB read + {k3['dram_write_bytes'] / 1e6:.2f} MB written per launch against 55.80 MB algorithmic (1024 x 54 488 B): {(k3['dram_read_bytes'] + k3['dram_write_bytes']) / 55795712:.2f} x, no re-reads; HBM fraction {r['hbm']['frac'] * 100:.1f} %.

Share of the step (ref-literal, from `r02_step_metrics.csv`, cold-cache serialised durations): K3 {dur['stwo_merkle_kernel']:.1f} us = {dur['stwo_merkle_kernel'] / tot * 100:.1f} %, K1 {dur['stwo_channel_ws_kernel']:.1f} us = {dur['stwo_channel_ws_kernel'] / tot * 100:.1f} % (latency-bound: 32 x 4 warps on
592 schedulers), K2 {dur['stwo_query_kernel']:.1f} us, K4 {dur['stwo_finalize_kernel']:.1f} us — the same shares as the CUDA-event `kernel_ms` of the bench line ({d['kernel_ms']['stwo_merkle']:.3f} / {d['kernel_ms']['stwo_channel']:.3f} / {d['kernel_ms']['stwo_query']:.3f} / {d['kernel_ms']['stwo_finalize']:.3f} ms).

Warp-state samples of K3 by opcode (source page of `{os.path.basename(merkle_rep)}`; {total_samples} samples): `SHF` {S('SHF', '# Samples')} (math-pipe throttle {S('SHF', 'stall_math')}, not selected {S('SHF', 'stall_not_selected')}), `LOP3` {S('LOP3', '# Samples')}
({S('LOP3', 'stall_math')} / {S('LOP3', 'stall_not_selected')}), `IMAD` {S('IMAD', '# Samples')} (dispatch {S('IMAD', 'stall_dispatch')}, wait {S('IMAD', 'stall_wait')}, selected {S('IMAD', 'stall_selected')}); long scoreboard (the sibling prefetch) {int(sum(v['stall_long_sb'] for v in op.values()))} in total, no-instruction {int(sum(v['stall_no_inst'] for v in op.values()))}: the kernel waits
for its two arithmetic pipes and for nothing else.
""")
sys.stdout.flush()
subprocess.run([sys.executable, os.path.join(HERE, "summarize_ncu.py"), merkle_rep, channel_rep])
