ncu --set full --clock-control none --import-source on -k regex:wit_lex -s 6 -c 1 -o gpurun_out/r01c_wit_lex python scratch/wit_time.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:wit_numbers -s 6 -c 1 -o gpurun_out/r01c_wit_numbers python scratch/wit_time.py > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r01c_launches.csv python bench.py --steps 4 --warmup 3 --pipeline 1 --no-cpu-baseline > /dev/null 2>&1
