set -x
mkdir -p gpurun_out
FHE_ONLY_NTT=1 timeout 300 python tools/microbench.py > gpurun_out/microbench_ntt.log 2>&1; grep fused gpurun_out/microbench_ntt.log | grep "1248\|312" | cut -c1-250
timeout 300 python tools/fused_prof.py | grep -E "rows=1248|per-group dur"
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_fused.json 2> gpurun_out/bench_fused.err; tail -c 1500 gpurun_out/bench_fused.json
FHE_NTT_FUSED=0 timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_twopass.json 2> gpurun_out/bench_twopass.err; tail -c 1500 gpurun_out/bench_twopass.json
