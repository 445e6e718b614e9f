set -x
mkdir -p gpurun_out
FHE_NTT_FUSED=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ntt_(fwd|inv)_pass" -c 4 -s 4 -o gpurun_out/prof_twopass -f python tools/profile_ntt.py > gpurun_out/ncu_twopass.log 2>&1; tail -2 gpurun_out/ncu_twopass.log
