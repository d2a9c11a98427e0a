# tuning sweep of the tabulated-sum commit (mult_kernels.cuh): batched-affine rounds x pairs per thread, cfg1, 34 GB table
for cfg in "5 0" "5 8" "5 12" "4 0" "4 12" "6 0"; do
  set -- $cfg
  if [ "$2" = "0" ]; then unset SBN_BA_BATCH; else export SBN_BA_BATCH=$2; fi
  SBN_MULT_ROUNDS=$1 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-prove --gens distinct 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('rounds=$1 batch=$2', 'value %.1fM'%(d['value']/1e6), 'ms %.3f'%d['ms_per_step'], 'e2e %.1fM'%(d['e2e']['value']/1e6))"
done
