set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_aes.py -m gpu -x -q -s -k "aes128" 2>&1 | grep -v Warning | tail -8
