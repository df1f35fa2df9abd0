set -u
M="--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv"
timeout 500 ncu $M --log-file gpurun_out/launches_r01s5_edsr_train.csv python profiles/ncu_target.py edsr_train > gpurun_out/ncu_edsr_train.log 2>&1
timeout 300 ncu $M --log-file gpurun_out/launches_r01s5_edsr_infer.csv python profiles/ncu_target.py edsr_infer > gpurun_out/ncu_edsr_infer.log 2>&1
timeout 700 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r01s5_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ls -la gpurun_out/*.csv | tail -5
python profiles/aggregate_launches.py gpurun_out/launches_r01s5_edsr_train.csv | head -30
