#!/bin/bash
# Full validation on one B200: GPU parity tests, smoke(), the default bench line.
#   gpurun --timeout 2400 -- 'bash tools/run_gpu_round.sh'
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; tail -3 gpurun_out/bench_r1.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1.json'))
for k in ('value','ms_per_step','gpu_launches','e2e','roofline','full_round','aes128','cpu_baseline','clocks'): print(k, d.get(k))
PY
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 | cut -c1-400
