# A/B runner: same bench, environment toggles given as arguments ("A=1 B=2" per arm), two runs per arm.
for arm in "$@"; do
echo "== $arm"
for i in 1 2; do
env $arm python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']/1e6,3), round(d['ms_per_step'],3), round(d['e2e']['value']/1e6,3), {k: round(v,3) for k,v in d['roofline']['per_kernel_ms_per_step'].items()}, d['clocks']['sm_mhz'])"
done
done
