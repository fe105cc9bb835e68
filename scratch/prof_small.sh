timeout 300 ncu --set full --clock-control none --import-source on -k "regex:logmel512|stats_pool|dense_finish|adam_kernel|pack_rows" -s 30 -c 12 -o gpurun_out/prof_small python scratch/one_step.py 4 2>&1 | tail -5
ls -la gpurun_out/
