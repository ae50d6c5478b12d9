# per-launch times of the fused ResBlock-round kernel under each tuning variant (ncu, one generator call at B = 8)
for v in "X=0" "PD_VOC_V64=1" "PD_VOC_V32=1" "PD_VOC_V32=2"; do
  n=$(echo $v | tr '=' '_')
  env $v ncu --kernel-name regex:voc_pair --metrics gpu__time_duration.sum --clock-control none -c 54 --csv --log-file gpurun_out/voc_ab_$n.csv python tools/voc_bench.py ${VB:-8} 1000 1 > /dev/null 2>&1
done
