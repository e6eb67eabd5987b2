SHORT="bench.py --steps 2 --warmup 3 --frames 128 --no-cpu --no-match --no-extra"
python $SHORT > gpurun_out/r2g_short.json 2> gpurun_out/r2g_short.err && \
ncu --set full --clock-control none --import-source on -k regex:k_ -s 42 -c 14 -f -o gpurun_out/r2g_full python $SHORT > gpurun_out/r2g_ncu.log 2>&1
ncu -i gpurun_out/r2g_full.ncu-rep --page raw --csv > gpurun_out/r2g_full_raw.csv 2>/dev/null
ls -la gpurun_out | tail -4
