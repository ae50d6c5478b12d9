# ncu cannot replay the cooperative cluster launch of the persistent stack kernel (the record comes back empty and the
# profiled process stops there), so the profiling passes drop the attribute with PD_COOP=0: same kernel, same grid.
set -x
export PD_COOP=0
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-text-front --no-wave-back"
$CMD > gpurun_out/r02f_bench_short_nocoop.json 2>/dev/null && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02f_launches.csv $CMD > /dev/null 2>&1
ncu --set full --clock-control none -k regex:layer_kernel -s 8 -c 1 -f -o /tmp/r02f_stack $CMD > /dev/null 2>&1
ncu -i /tmp/r02f_stack.ncu-rep --page raw --csv > gpurun_out/r02f_stack_raw.csv 2>/dev/null
rm -f /tmp/r02f_stack.ncu-rep
