timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 93 -c 62 --csv --log-file gpurun_out/r1b_launches_warm.csv python scratch/one_step.py 5 > /dev/null 2>&1
tail -3 gpurun_out/r1b_launches_warm.csv | cut -c1-200
