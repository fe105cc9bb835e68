for v in v0_current v1_bias_global v2_no_mask_prefetch v3_10warps v4_old_like; do
  echo "=== $v"
  LBX_LIB=$PWD/scratch/libs/$v.so timeout 200 python scratch/gemm_bench.py 2>&1 | grep "tile_n=256" | grep -E "frame1 fwd|frame2 fwd|frame4 fwd|frame3 dgrad|frame2 dgrad p1|big square|frame2 wgrad"
done
