set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_aes.py -m gpu -x -q -s 2>&1 | grep -v Warning | tail -8
timeout 600 python tools/boot_phases.py 12 2>&1 | grep -v Warning | tail -10
