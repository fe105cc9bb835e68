export LBX_GEMM_PAIR=1
timeout 120 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x 2>&1 | tail -6
timeout 120 python scratch/gemm_bench.py 2>&1 | grep "tile_n=256" | head -20
