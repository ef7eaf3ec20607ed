mkdir -p gpurun_out
timeout 300 python tools/prof_target.py conv3d 3 > gpurun_out/p0.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stem_t -s 1 -c 1 -o gpurun_out/r02f_stem_t -f python tools/prof_target.py conv3d 3 > gpurun_out/ncu_stem.log 2>&1; echo "ncu stem rc=$?"
timeout 300 python tools/prof_target.py forward 3 > gpurun_out/p1.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02f_launches_forward.csv python tools/prof_target.py forward 3 > gpurun_out/ncu_fwd.log 2>&1; echo "ncu launches rc=$?"
ls -la gpurun_out/*.ncu-rep
