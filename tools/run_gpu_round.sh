#!/bin/bash
# Full validation on one B200: GPU parity tests, smoke(), the default bench line (both arms), then the ncu launch list of
# a short bench run and ncu --set full captures of the top kernels.
#   gpurun --timeout 3000 -- 'bash tools/run_gpu_round.sh [tag]'
tag=${1:-r02}
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -3 gpurun_out/bench_$tag.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$tag.json'))
for k in ('value','ms_per_step','ms_per_round','gpu_launches','e2e','stage_ms','roofline','roofline_fp64','roofline_keyswitch_inner','subbytes','cpu_baseline','clocks','ntt_rows_per_state','hbm_peak_allocated_gb'): print(k, d.get(k))
print(open('gpurun_out/bench_ref_$tag.json').read()[:600])
PY
# launch list (cold-cache, serialised: shares only) of one AES-128 pass; the same command ran plain just above
short="python bench.py --steps 1 --warmup 1 --states 1 --no-extras --no-cpu-baseline"
timeout 900 $short > gpurun_out/plain_short.log 2>&1 &&
timeout 2400 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_$tag.csv $short > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log
