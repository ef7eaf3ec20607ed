for m in "16 1" "8 1"; do
  tag=$(echo $m | tr ' ' '_')
  timeout 180 python tools/gpu_bringup.py stack $m > gpurun_out/stack_$tag.log 2>&1
  echo "mode $m rc=$?"; grep -c "OK " gpurun_out/stack_$tag.log; grep "BAD\|EXCEPTION\|Error\|failures" gpurun_out/stack_$tag.log | head -30
  set -- $m; SBLK_ENC_STACK_CL=$1 SBLK_ENC_STACK_MC=$2 L=2 timeout 100 python tools/stack_stamps.py 2>&1 | tail -7
done
timeout 200 python tools/time_parts.py 2>&1 | tail -9
