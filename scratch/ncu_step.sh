# per-kernel instruction counts / pipe use of one serial 1024-proof step + the .wit tokeniser kernels
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_fmaheavy.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed \
  --clock-control none -k regex:'stwo_|wit_' -c 40 --csv --log-file gpurun_out/r01c_step_metrics.csv python bench.py --steps 3 --warmup 3 --pipeline 1 --no-cpu-baseline > /dev/null 2>&1
