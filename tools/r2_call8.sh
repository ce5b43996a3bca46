#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/r2c8_pytest.log 2>&1; echo "pytest exit $?" >> $out/r2c8_pytest.log
tail -8 $out/r2c8_pytest.log | cut -c1-300
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ctc_warp -s 1 -c 1 -o $out/r2_ctc_warp python tools/profile_target.py ctc > $out/r2c8_ncu_ctc.log 2>&1; tail -2 $out/r2c8_ncu_ctc.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gated_fwd -s 1 -c 1 -o $out/r2_lstm_fwd python tools/profile_target.py lstm 200 > $out/r2c8_ncu_lstm.log 2>&1; tail -2 $out/r2c8_ncu_lstm.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gated_bwd -s 1 -c 1 -o $out/r2_lstm_bwd python tools/profile_target.py lstm 200 > $out/r2c8_ncu_lstmb.log 2>&1; tail -2 $out/r2c8_ncu_lstmb.log
