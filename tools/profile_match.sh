#!/bin/bash
# tools/profile_match.sh TAG -- on the GPU box: the matching kernels (C5 kNN-2 / SearchByProjection, C2 / C3 batched
# stereo matchers) under ncu --set full, and a C4 (1280x720) capture of the extraction kernels for traffic.json.
# The plain commands run first; ncu only after they exited 0.  One ncu pass per kernel family (few launches each), raw
# CSVs concatenated; the .ncu-rep files are dropped (gpurun_out/ carries at most 64 MiB back).
T=${1:-rX}
O=gpurun_out
mkdir -p $O
python tools/match_driver.py 3 > $O/${T}_match_plain.log 2>&1 || { echo "driver failed"; tail -5 $O/${T}_match_plain.log; exit 1; }
tail -4 $O/${T}_match_plain.log
i=0
for fam in 'k_knn2_umma|k_knn2_merge' 'k_search|k_build_grid|k_claims|k_assign' 'k_stereo'; do
  i=$((i+1))
  ncu --set full --clock-control none -k "regex:$fam" -c 8 -f -o $O/${T}_match$i python tools/match_driver.py 1 > $O/${T}_ncu_match$i.log 2>&1
  ncu -i $O/${T}_match$i.ncu-rep --page raw --csv > $O/${T}_match_raw$i.csv 2>/dev/null
  rm -f $O/${T}_match$i.ncu-rep
done
if [ "$2" != "nomatchonly" ]; then
C4="bench.py --workload c4 --steps 2 --warmup 3 --frames 64 --no-cpu --no-match --no-extra"
python $C4 > $O/${T}_c4short.json 2> $O/${T}_c4short.err || { echo "c4 short failed"; exit 1; }
ncu --set full --clock-control none -k regex:k_ -s 42 -c 14 -f -o $O/${T}_c4full python $C4 > $O/${T}_ncu_c4.log 2>&1
ncu -i $O/${T}_c4full.ncu-rep --page raw --csv > $O/${T}_c4full_raw.csv 2>/dev/null
rm -f $O/${T}_c4full.ncu-rep
fi
du -sh $O
ls -la $O | tail -8
