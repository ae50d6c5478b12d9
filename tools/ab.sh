for i in 1 2; do
for f in 0 2 4; do
PD_LAYER_FLAGS=$f python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('flags=$f', round(d['value']/1e6,3), round(d['ms_per_step'],3), round(d['roofline']['avg_launch_us'],2), d['clocks']['sm_mhz'])"
done; done
