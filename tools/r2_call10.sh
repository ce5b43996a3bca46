#!/bin/bash
out=gpurun_out; mkdir -p $out
run() { env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$*', round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['kernel_ms_per_step'].items()})"; }
run A=1
run CTCASR_LSTM_L2_KEEP_MB=90
run CTCASR_LSTM_L2_KEEP_MB=110
run CTCASR_LSTM_L2_KEEP_MB=40
run CTCASR_LSTM_STAGGER_NS=0
run CTCASR_LSTM_STAGGER_NS=6000
timeout 600 ncu --set full --clock-control none -k regex:gated_fwd -s 1 -c 1 -o $out/r2_lstm_fwd_T1000 python tools/profile_target.py lstm 1000 > $out/r2c10_ncu1.log 2>&1; tail -1 $out/r2c10_ncu1.log
timeout 600 ncu --set full --clock-control none -k regex:gated_bwd -s 1 -c 1 -o $out/r2_lstm_bwd_T1000 python tools/profile_target.py lstm 1000 > $out/r2c10_ncu2.log 2>&1; tail -1 $out/r2c10_ncu2.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $out/r2_ncu_launches_cfg2_step.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/r2c10_ncu_bench.log 2>&1; tail -2 $out/r2_ncu_launches_cfg2_step.csv | cut -c1-200
