timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for mode in ref-literal prover-consistent; do
  echo "== $mode DEDUP 1 pipelined 1024"
  python bench.py --no-cpu-baseline --steps 1000 --mode $mode 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['kernel_ms'], d['gpu_launches'])"
done
bash scratch/ncu_dd.sh prover-consistent | grep -A1 "check" | grep duration
