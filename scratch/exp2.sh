timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for mode in ref-literal prover-consistent; do
for d in 1 0; do
  echo "== $mode DEDUP $d pipelined 1024"
  SSYM_MERKLE_DEDUP=$d python bench.py --no-cpu-baseline --steps 500 --mode $mode 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['kernel_ms'], d['gpu_launches'])"
  echo "== $mode DEDUP $d serial 8192"
  SSYM_MERKLE_DEDUP=$d python bench.py --no-cpu-baseline --batch 8192 --pipeline 1 --copies 1 --steps 20 --warmup 3 --mode $mode 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['kernel_ms'])"
done
done
