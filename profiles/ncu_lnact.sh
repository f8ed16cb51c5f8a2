# ncu evidence for the current build: launch list of the bench step, one full capture each of the fused LayerNorm + tanh kernels (C3 shapes)
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --lite --steps 2 --warmup 1 > gpurun_out/ncu_launches.log 2>&1
for k in fwd bwd; do
  ncu --set full --import-source on --clock-control none -k regex:lnact_feat_${k} --launch-skip 7 --launch-count 1 -o /tmp/lnact_$k -f python profiles/c3_probe.py c3 > gpurun_out/ncu_lnact_$k.log 2>&1
  ncu -i /tmp/lnact_$k.ncu-rep --page raw --csv > gpurun_out/r02_full_lnact_${k}_raw.csv 2>/dev/null
  ncu -i /tmp/lnact_$k.ncu-rep --page source --csv --print-source sass > gpurun_out/r02_lnact_${k}_source.csv 2>/dev/null
done
ls -la gpurun_out/r02_*lnact* gpurun_out/r02_launches.csv
