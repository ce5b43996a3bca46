#!/bin/bash
# ncu --set full captures of the current kernels (one GPU, a few launches each); reports land in gpurun_out/.
# Usage (on the box): bash tools/ncu_round.sh <tag>
tag=${1:-x}
out=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:gemm_tc -s 3 -c 3 -o $out/${tag}_gemm python tools/profile_target.py gemm > $out/${tag}_ncu_gemm.log 2>&1
timeout 300 $NCU -k regex:gemm_tc -s 3 -c 3 -o $out/${tag}_convgemm python tools/profile_target.py convgemm > $out/${tag}_ncu_convgemm.log 2>&1
timeout 300 $NCU -k regex:"im2col|col2im" -s 3 -c 3 -o $out/${tag}_conv python tools/profile_target.py conv > $out/${tag}_ncu_conv.log 2>&1
timeout 300 $NCU -k regex:ctc_loss -s 1 -c 1 -o $out/${tag}_ctc python tools/profile_target.py ctc > $out/${tag}_ncu_ctc.log 2>&1
timeout 300 $NCU -k regex:beam_search -s 1 -c 1 -o $out/${tag}_beam python tools/profile_target.py beam > $out/${tag}_ncu_beam.log 2>&1
ls -la $out/${tag}_*.ncu-rep
