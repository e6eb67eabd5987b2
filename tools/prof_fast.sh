SHORT="bench.py --steps 2 --warmup 3 --frames 128 --no-cpu --no-match"
python $SHORT > gpurun_out/r2b_short.json 2> gpurun_out/r2b_short.err && \
ncu --set full --clock-control none --import-source on -k regex:k_fast_cells -s 3 -c 1 -f -o gpurun_out/r2b_fast python $SHORT > gpurun_out/r2b_ncu.log 2>&1
ls -la gpurun_out | tail -5
