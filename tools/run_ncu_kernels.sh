#!/bin/bash
# ncu --set full captures of the hot kernels inside the bit bootstrap of one AES-128 pass: CoeffToSlot (ModUp, fused baby steps,
# giant-step key switches) and EvalMod (fused multiply + relinearise + rescale).  The probe runs plain first; ncu runs on the
# same command line after it exited 0.  The reports are turned into raw-page CSVs on the box (a report with source is ~5 MB per
# kernel; gpurun brings back at most 64 MiB).
#   gpurun --timeout 2400 -- 'bash tools/run_ncu_kernels.sh r02'
tag=${1:-r02}
mkdir -p gpurun_out
cmd="python tools/aes_bits_probe.py --states 1 --rounds 4 --warmup 1 --nvtx"
timeout 600 $cmd > gpurun_out/ncu_plain.log 2>&1 || { echo plain run failed; tail -5 gpurun_out/ncu_plain.log; exit 1; }
timeout 1500 ncu --set full --clock-control none --nvtx --nvtx-include "after_boot_mod_raise/" \
   -k regex:"ntt_fwd_chained|ntt_inv_chained|k_bconv|k_bsgs_inner|k_ks_inner" -c 14 \
   -o /tmp/ncu_cts_$tag -f $cmd > gpurun_out/ncu_cts.log 2>&1
tail -2 gpurun_out/ncu_cts.log
ncu -i /tmp/ncu_cts_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_cts_$tag.csv 2>/dev/null
timeout 1500 ncu --set full --clock-control none --nvtx --nvtx-include "after_boot_conjugate_split/" \
   -k regex:"ntt_fwd_chained|ntt_inv_chained|k_bconv|k_ks_inner|k_lincomb" -c 10 \
   -o /tmp/ncu_evalmod_$tag -f $cmd > gpurun_out/ncu_evalmod.log 2>&1
tail -2 gpurun_out/ncu_evalmod.log
ncu -i /tmp/ncu_evalmod_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_evalmod_$tag.csv 2>/dev/null
ls -la gpurun_out/
