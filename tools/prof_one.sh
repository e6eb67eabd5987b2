# tools/prof_one.sh KERNEL_REGEX TAG [SKIP]: one `ncu --set full` capture of one kernel of the short bench (after the plain run exited 0)
SHORT="bench.py --steps 2 --warmup 3 --frames 128 --no-cpu --no-match --no-extra"
python $SHORT > gpurun_out/$2_short.json 2> gpurun_out/$2_short.err && \
ncu --set full --clock-control none --import-source on -k regex:$1 -s ${3:-3} -c 1 -f -o gpurun_out/$2 python $SHORT > gpurun_out/$2_ncu.log 2>&1
ncu -i gpurun_out/$2.ncu-rep --page raw --csv > gpurun_out/$2_raw.csv 2>/dev/null
ncu -i gpurun_out/$2.ncu-rep --page source --csv > gpurun_out/$2_src.csv 2>/dev/null
ls -la gpurun_out/$2*
