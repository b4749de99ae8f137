# r01e: launch list of the bench command, full captures of the dominant kernel and of the compact-form expansion
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01e_launches.csv python bench.py --steps 4 --warmup 3 --pipeline 1 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:stwo_merkle_kernel -s 3 -c 1 -o gpurun_out/r01e_merkle python bench.py --steps 4 --warmup 3 --pipeline 1 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:stwo_expand_kernel -s 4 -c 1 -o gpurun_out/r01e_expand python bench.py --steps 4 --warmup 3 --pipeline 1 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/ | tail -5
