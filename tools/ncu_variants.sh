mkdir -p gpurun_out/r01i
for v in "default:" "wl2:CONZIC_WIDE_LN=2" "wl1m2:CONZIC_WIDE_LN=1 CONZIC_WIDE_LN_MODE=2"; do
  tag=${v%%:*}; envs=${v#*:}
  env $envs timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 182 -c 400 --csv --log-file gpurun_out/r01i/launches_$tag.csv python tools/profile_step.py --ii 0 --steps 1 --warm 1 > gpurun_out/r01i/ncu_$tag.log 2>&1; echo "$tag rc=$?"
done
