for i in 1 2; do
for f in 1 0; do
PD_STACK=$f python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('stack=$f', round(d['value']/1e6,3), round(d['ms_per_step'],3), round(d['roofline']['avg_launch_us'],2), d['roofline']['launches'], d['clocks']['sm_mhz'], d['gpu_launches'])"
done; done
