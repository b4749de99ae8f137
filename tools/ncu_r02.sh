#!/usr/bin/env bash
# ncu captures behind profiles/r02_* (run under gpurun on ONE GPU; numbers printed by a run under ncu are never bench values).
#   gpurun --timeout 1500 -- 'bash tools/ncu_r02.sh'
# then here: python profiles/pipe_counts.py gpurun_out/r02_step_metrics.csv gpurun_out/r02_shared_step_metrics.csv --proofs 1024,1024
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
M=gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fmaheavy.sum,dram__bytes_read.sum,dram__bytes_write.sum
HEAD="python bench.py --steps 1 --warmup 3 --passes 4 --pipeline 1 --headline-only"
# per-kernel instruction / pipe / DRAM counts of one serial step, both semantics (1024 proofs per launch)
ncu --metrics $M --clock-control none -k regex:stwo_ -s 24 -c 12 --csv --log-file gpurun_out/r02_step_metrics.csv $HEAD > /dev/null 2> gpurun_out/r02_step_metrics.err
ncu --metrics $M --clock-control none -k regex:stwo_ -s 40 -c 16 --csv --log-file gpurun_out/r02_shared_step_metrics.csv $HEAD --mode prover-consistent > /dev/null 2> gpurun_out/r02_shared_step_metrics.err
# launch list of the bench command (shares, not absolutes: cold caches, serialised)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --passes 8 --e2e-passes 4 --no-cpu-baseline > /dev/null 2> gpurun_out/r02_launches.err
# full captures of the dominant kernel and of the latency-bound transcript kernel
ncu --set full --clock-control none --import-source on -k regex:stwo_merkle_kernel -s 8 -c 1 -f -o gpurun_out/r02_merkle $HEAD > /dev/null 2> gpurun_out/r02_merkle.err
ncu --set full --clock-control none --import-source on -k regex:stwo_channel -s 8 -c 1 -f -o gpurun_out/r02_channel $HEAD > /dev/null 2> gpurun_out/r02_channel.err
ls -la gpurun_out/r02_*
