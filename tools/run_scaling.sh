#!/bin/bash
# One point of the scaling tables: bash tools/run_scaling.sh <n_gpus> <weak|strong> [total_states]
#   gpurun --gpus N --timeout 1200 -- 'bash tools/run_scaling.sh N strong 8'
n=$1; mode=$2; tot=${3:-8}
mkdir -p gpurun_out
out=gpurun_out/scale_${mode}_${n}.json
args="--gpus $n --steps 3 --warmup 3 --no-extras --no-cpu-baseline --scaling $mode --total-states $tot"
if [ "$n" = "1" ]; then
  timeout 1000 python bench.py $args > $out 2> gpurun_out/scale_${mode}_${n}.err
else
  timeout 1000 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 bench.py $args > $out 2> gpurun_out/scale_${mode}_${n}.err
fi
echo rc=$?; tail -2 gpurun_out/scale_${mode}_${n}.err
python - <<PY
import json
l=[x for x in open("$out").read().split("\n") if x.startswith("{")]
d=json.loads(l[-1])
print({k:d.get(k) for k in ("n_gpus","scaling","value","ms_per_step","e2e","startup","clocks")}, d["config"]["states_per_gpu"])
PY
