mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/launches_tmp.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-pass1 > gpurun_out/ncu_list.log 2>&1
python profiles/summarize.py launches gpurun_out/launches_tmp.csv
