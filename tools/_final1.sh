set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2f_gputest.txt
SCZ_MSM_STREAM=1 python tools/hp_time.py 20 4 2>&1 | grep -E "rep 3" > gpurun_out/r2f_hp_time.txt
python bench.py --steps 20 > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err
SCZ_MSM_STREAM=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_hp_ncu_launches.csv python tools/hp_time.py 20 1 > /dev/null 2>&1
cat gpurun_out/r2f_gputest.txt gpurun_out/r2f_hp_time.txt
