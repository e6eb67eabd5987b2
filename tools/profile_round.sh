#!/bin/bash
# tools/profile_round.sh TAG -- run on the GPU box (under gpurun): GPU test suite, default bench (both arms),
# then -- only after the plain runs exited 0 -- the ncu launch list and one `--set full` capture of a short
# bench with the same kernels.  Everything lands in gpurun_out/TAG_*; summaries are copied to profiles/ by hand.
T=${1:-rX}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/${T}_pytest.log
tail -3 $O/${T}_pytest.log
python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err || { echo "bench failed"; tail -5 $O/${T}_bench.err; exit 1; }
cat $O/${T}_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_bench_ref.json 2> $O/${T}_bench_ref.err; cat $O/${T}_bench_ref.json
SHORT="bench.py --steps 2 --warmup 3 --frames 128 --no-cpu --no-match --no-extra"
python $SHORT > $O/${T}_short.json 2> $O/${T}_short.err || { echo "short bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${T}_launches.csv \
    python $SHORT > $O/${T}_ncu_launch.log 2>&1
# launch list of the bench command at its own batch size (first five steps = 70 launches of 1024 frames)
ncu --metrics gpu__time_duration.sum --clock-control none -c 140 --csv --log-file $O/${T}_launches_1024.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-match --no-extra > $O/${T}_ncu_launch_1024.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ -s 42 -c 14 -f -o $O/${T}_full \
    python $SHORT > $O/${T}_ncu_full.log 2>&1
ncu -i $O/${T}_full.ncu-rep --page raw --csv > $O/${T}_full_raw.csv 2>/dev/null
ls -la $O | tail -12
