#!/bin/bash
out=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 400 $NCU -k regex:"conv_fwd|conv_wgrad|conv_dgrad" -s 3 -c 3 -o $out/r2c32_conv_implicit python tools/profile_target.py conv > $out/r2c32_ncu_conv.log 2>&1; tail -1 $out/r2c32_ncu_conv.log
