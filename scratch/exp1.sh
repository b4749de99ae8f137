./scratch/pipe_probe
for m in 1 6 7; do
  echo "== ADDMODE $m pipelined 1024"
  SSYM_ADDMODE=$m python bench.py --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['kernel_ms'])"
  echo "== ADDMODE $m serial 16384"
  SSYM_ADDMODE=$m python bench.py --no-cpu-baseline --batch 16384 --pipeline 1 --copies 1 --steps 20 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['kernel_ms'])"
done
