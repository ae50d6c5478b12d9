# End-of-round measurement pass (one gpurun call): bench line, launch list of the same command, ncu --set full of the dominant
# sampler kernel and of the vocoder's convolution kernels (exported to CSV on the box: the .ncu-rep files are too big to bring
# back).  Outputs under gpurun_out/r02f_*.
set -x
python bench.py > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-text-front --no-wave-back"
$CMD > gpurun_out/r02f_bench_short.json 2>/dev/null && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02f_launches.csv $CMD > /dev/null 2>&1
cap() {  # name, kernel regex, skip, then the command
  local name=$1 k=$2 s=$3; shift 3
  ncu --set full --clock-control none -k regex:$k -s $s -c 1 -f -o /tmp/$name "$@" > /dev/null 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  rm -f /tmp/$name.ncu-rep
}
cap r02f_stack layer_kernel 8 $CMD
VOC="python tools/voc_bench.py 32 1000 1"
$VOC
cap r02f_voc_conv256 voc_conv_kernel 51 $VOC
cap r02f_voc_conv128 voc_conv_kernel 70 $VOC
cap r02f_voc_pair64 voc_pair_kernel 30 $VOC
python tools/voc_bench.py 64 1000 3; python tools/voc_bench.py 8 1000 5; python tools/voc_bench.py 1 1000 10
ls -la gpurun_out/
