for i in 1 2; do
for f in 1 0; do
PD_IDENT_RESIDENT=$f python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ident_res=$f', round(d['value']/1e6,3), round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['roofline']['per_kernel_ms_per_step'].items()}, d['clocks']['sm_mhz'])"
done; done
