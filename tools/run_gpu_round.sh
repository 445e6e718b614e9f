set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "keyswitch or engine_ops" 2>&1 | tail -3
timeout 600 python tools/microbench.py > gpurun_out/microbench.log 2>&1; grep "ks\.\|keyswitch" gpurun_out/microbench.log | grep '"nq": 31\|"nq": 21' | cut -c1-220
