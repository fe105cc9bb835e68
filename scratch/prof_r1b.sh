timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 93 -c 40 --csv --log-file gpurun_out/r1b_launches_train_step.csv python scratch/one_step.py 4 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -s 93 -c 31 -o gpurun_out/r1b_full_step python scratch/one_step.py 4 > /dev/null 2>&1
ls -la gpurun_out | grep r1b
