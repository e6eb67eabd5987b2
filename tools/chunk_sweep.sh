#!/bin/bash
# tools/chunk_sweep.sh -- e2e throughput of the host-pointer batch API versus the pipeline chunk size (frames per chunk).
python -m pytest tests/test_gpu_match.py tests/test_gpu_extract.py tests/test_golden.py -x -q -m gpu -k "distinctive or extract or golden" 2>&1 | tail -2
for c in 64 128 192 256 384 512; do
  python bench.py --steps 10 --warmup 3 --no-cpu --no-match --chunk $c > gpurun_out/sweep_$c.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/sweep_$c.json')); print('chunk', $c, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'fast_score', round(d['roofline']['kernel_ms_per_step']['fast_score'],3))"
done
