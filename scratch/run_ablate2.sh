timeout 200 python -m pytest tests/test_xvector_gpu.py -m gpu -q -x 2>&1 | tail -3
for l in base "" u4s1 u2s0 u1s1; do
  if [ -z "$l" ]; then timeout 120 python scratch/ablate2.py 2>&1 | tail -1; else LBX_LIB=$PWD/scratch/libs/lib_$l.so timeout 120 python scratch/ablate2.py 2>&1 | tail -1; fi
done
