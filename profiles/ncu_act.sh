for k in fwd bwd; do
  ncu --set full --import-source on --clock-control none -k regex:^act_${k}_kernel --launch-skip 3 --launch-count 1 -o /tmp/act_$k -f python profiles/c3_probe.py c4 > gpurun_out/ncu_act_$k.log 2>&1
  ncu -i /tmp/act_$k.ncu-rep --page raw --csv > gpurun_out/r02z_full_act_${k}_raw.csv 2>/dev/null
done
ls -la gpurun_out/r02z_full_act*; tail -2 gpurun_out/ncu_act_fwd.log
