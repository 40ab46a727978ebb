timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k 'leading_dimension or shorter_than' 2>&1 | tail -5
for sl in 1024 1536 2048 2560; do
  timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-workloads --slab $sl 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('slab $sl', 'e2e', round(d['e2e']['value']), {k: round(v['value']) for k,v in d['e2e']['variants'].items()})"
done
TAG=pf VARIANTS='A=1 ATTRICI_PREP_PREFETCH=0' bash scripts/gpu_ab_env.sh
