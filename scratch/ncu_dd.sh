MODE=${1:-prover-consistent}
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:stwo_ -s 16 -c 8 --csv --log-file gpurun_out/dd.csv python bench.py --no-cpu-baseline --batch 8192 --pipeline 1 --copies 1 --steps 3 --warmup 3 --mode $MODE > /dev/null 2>&1; python - <<EOF
import csv
rows=list(csv.reader(open("gpurun_out/dd.csv")))
hi=[i for i,r in enumerate(rows) if "Kernel Name" in r][0]
h=rows[hi]
for r in rows[hi+1:]:
    if len(r)>len(h)-1: print(r[h.index("Kernel Name")][:50], r[h.index("Metric Name")], r[h.index("Metric Value")])
EOF
