# A/B of the fused-round kernel's tuning variants: whole-generator time at two batch sizes
for v in "X=0" "PD_VOC_V16=1" "PD_VOC_V16=2" "PD_VOC_V32=1" "PD_VOC_V32=2"; do
  echo "== $v"
  env $v python tools/voc_bench.py 64 1000 3
  env $v python tools/voc_bench.py 8 1000 5
done
