#!/bin/bash
# tools/fast_per_sweep.sh -- k_fast_score time versus tiles per CTA (ORBFE_FAST_TILES_PER_CTA overrides the heuristic).
for p in 8 12 16 24 32; do ORBFE_FAST_TILES_PER_CTA=$p python bench.py --steps 10 --warmup 3 --no-cpu --no-match 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print($p, round(d['value']), round(d['e2e']['value']), round(d['roofline']['kernel_ms_per_step']['fast_score'],3))"; done
