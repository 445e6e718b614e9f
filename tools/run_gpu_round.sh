set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python tools/microbench.py > gpurun_out/microbench.log 2>&1; grep "two_pass\|rescale\|keyswitch\|ks\." gpurun_out/microbench.log | grep -v '"nq": 11\|"nq": 21' | cut -c1-220
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_new.json 2> gpurun_out/bench_new.err; python -c "
import json; d=json.load(open('gpurun_out/bench_new.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches','ms_per_ciphertext')}, d['e2e'], d['roofline'], d['full_round'])"
