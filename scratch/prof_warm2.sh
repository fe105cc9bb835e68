timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 84 -c 56 --csv --log-file gpurun_out/r1c_launches_warm.csv python scratch/one_step.py 5 > /dev/null 2>&1
