mkdir -p gpurun_out
for v in 3072 5120 4096 3072 5120; do
  E2B_RT_EW8_MAXK=$v timeout 300 python bench.py --no-cpu-baseline --steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('E2B_RT_EW8_MAXK=$v', round(d['value'],2), d['clocks']['sm_mhz'], {k:v['ms'] for k,v in d['forward']['by_kind'].items() if k=='gemm_resid'})"
done | tee gpurun_out/r2_ew8_sweep31.txt
