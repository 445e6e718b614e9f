set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for b in 8 16; do timeout 900 python bench.py --steps 3 --warmup 3 --batch $b --no-cpu-baseline --no-full-round --no-aes128 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('batch', d['config']['batch_ciphertexts_per_gpu'], 'value', round(d['value']), 'ms/ct', round(d['ms_per_ciphertext'],3), 'e2e', round(d['e2e']['value']))"; done
