python -m pytest tests/test_gpu_vocoder.py -m gpu -x -q 2>&1 | tail -2
for v in "X=0" "PRODIFF_B200_LIB=/root/repo/prodiff_b200/csrc/libprodiff_b200_s200.so"; do
  echo "== $v"
  env $v python tools/voc_bench.py 64 1000 3
  env $v python tools/voc_bench.py 8 1000 5
done
PRODIFF_B200_LIB=/root/repo/prodiff_b200/csrc/libprodiff_b200_s200.so python -m pytest tests/test_gpu_vocoder.py -m gpu -x -q 2>&1 | tail -2
python tools/voc_pair_trace.py 16 3 1 8 512000 2>&1 | tail -3
