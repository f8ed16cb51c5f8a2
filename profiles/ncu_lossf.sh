set -x
ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k 'regex:int.128, .bool.1>' --launch-skip 2 --launch-count 1 -o /tmp/lossf -f python bench.py --lite --steps 2 --warmup 1 > gpurun_out/ncu_lossf.log 2>&1
ncu -i /tmp/lossf.ncu-rep --page raw --csv > gpurun_out/r02z_full_lossf_raw.csv 2>/dev/null
ncu -i /tmp/lossf.ncu-rep --page source --csv --print-source sass > gpurun_out/r02z_lossf_source.csv 2>/dev/null
ls -la gpurun_out/r02z_*; tail -3 gpurun_out/ncu_lossf.log
