set -x
python bench.py > gpurun_out/bench_r1m.json 2> gpurun_out/bench_r1m.err
python tools/prof_target.py forward 2 > gpurun_out/plain_fwd.log 2>&1 && ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1m.csv python tools/prof_target.py forward 2 > gpurun_out/ncu_fwd.log 2>&1
python tools/prof_target.py stack 3 > gpurun_out/plain_stack.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:encoder_stack -s 1 -c 1 -f -o gpurun_out/prof_stack_r1b python tools/prof_target.py stack 3 > gpurun_out/ncu_stack.log 2>&1
tail -2 gpurun_out/ncu_stack.log
