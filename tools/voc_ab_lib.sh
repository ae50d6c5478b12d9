# same-box A/B of two builds (libprodiff_b200_prev.so vs libprodiff_b200.so): ncu per-launch times of the convolution kernels, 2 rounds
for r in 1 2; do for v in prev cur; do
  L=/root/repo/prodiff_b200/csrc/libprodiff_b200.so; [ $v = prev ] && L=/root/repo/prodiff_b200/csrc/libprodiff_b200_prev.so
  PRODIFF_B200_LIB=$L ncu --metrics gpu__time_duration.sum --clock-control none -k regex:voc_conv_kernel -s 43 -c 43 --csv --log-file gpurun_out/ab_${v}_$r.csv python tools/voc_bench.py 32 1000 1 > /dev/null 2>&1
done; done
for r in 1 2; do for v in prev cur; do
  L=/root/repo/prodiff_b200/csrc/libprodiff_b200.so; [ $v = prev ] && L=/root/repo/prodiff_b200/csrc/libprodiff_b200_prev.so
  echo $v; PRODIFF_B200_LIB=$L python tools/voc_bench.py 64 1000 3
done; done
