for v in g2 g1; do
  echo "=== $v"
  LBX_LIB=$PWD/scratch/libs/$v.so timeout 200 python scratch/gemm_bench.py 2>&1 | grep "tile_n=256" | grep -E "frame. fwd|dgrad|big square|frame2 wgrad"
done
LBX_LIB=$PWD/scratch/libs/g2.so timeout 300 python -m pytest tests/test_gemm_gpu.py tests/test_xvector_gpu.py -m gpu -q -x 2>&1 | tail -3
