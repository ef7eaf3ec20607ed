set -x
ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__m_xbar2l1tex_read_bytes.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1i.csv python tools/prof_target.py forward 2 > gpurun_out/ncu_fwd.log 2>&1
for w in l3conv l4conv; do
  ncu --set full --clock-control none --import-source on -k regex:igemm2 -s 1 -c 1 -f -o gpurun_out/prof_${w}_r1c python tools/prof_target.py $w 3 > gpurun_out/ncu_$w.log 2>&1
done
