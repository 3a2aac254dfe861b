mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "not full_size and not 2048 and not 3d and not config_" 2>&1 | tail -3
for rep in 1 2; do
for v in old new; do
  cp tools/_ab/libleniax_b200_$v.so leniax_b200/libleniax_b200.so
  echo "== $v"
  python bench.py --no-secondary --no-cpu-baseline --steps 6 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('B', d['value'], d['ms_per_step'], d['roofline']['frac'])"
  python bench.py --config C --no-secondary --no-cpu-baseline --steps 3 --warmup 2 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C', d['value'], d['ms_per_step'], d['roofline']['frac'])"
done; done
