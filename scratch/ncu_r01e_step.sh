# r01e: per-kernel instruction counts / pipe use of serial 1024-proof steps (ref-literal loop, then the prover-consistent leg with the shared-node schedule)
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fmaheavy.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none -k regex:'stwo_' -s 12 -c 12 --csv --log-file gpurun_out/r01e_step_metrics.csv python bench.py --steps 3 --warmup 3 --pipeline 1 --no-cpu-baseline > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fmaheavy.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none -k regex:'stwo_' -s 16 -c 8 --csv --log-file gpurun_out/r01e_shared_step_metrics.csv python bench.py --no-cpu-baseline --batch 8192 --pipeline 1 --copies 1 --steps 3 --warmup 3 --mode prover-consistent > /dev/null 2>&1
