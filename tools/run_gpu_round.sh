set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; tail -3 gpurun_out/bench_r1.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1.json'))
for k in ('value','ms_per_step','gpu_launches','e2e','roofline','full_round','aes128','cpu_baseline','clocks'): print(k, d.get(k))
PY
