export LBX_GEMM_PAIR=1
timeout 200 python -m pytest tests/test_xvector_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_pair.json
python -c "
import json; d=json.load(open('gpurun_out/bench_pair.json')); print('PAIR', d['ms_per_step'], d['value'], d['roofline']['gemm_ms_per_step'], d['roofline']['frac'], d['e2e']['ms_per_step'])"
export LBX_GEMM_PAIR=0
timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_nopair.json
python -c "
import json; d=json.load(open('gpurun_out/bench_nopair.json')); print('SINGLE', d['ms_per_step'], d['value'], d['roofline']['gemm_ms_per_step'], d['roofline']['frac'], d['e2e']['ms_per_step'])"
