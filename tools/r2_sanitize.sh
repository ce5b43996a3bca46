#!/bin/bash
out=gpurun_out; mkdir -p $out
CS=/usr/local/cuda/bin/compute-sanitizer
run() { tool=$1; shift; name=$1; shift; timeout 900 $CS --tool $tool --print-limit 5 python -m pytest "$@" -x -q > $out/r2san_${tool}_${name}.log 2>&1; echo "$tool $name: exit $? :: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' $out/r2san_${tool}_${name}.log | tr '\n' ' ')"; }
run memcheck rnn "tests/test_gpu_recurrence.py::test_one_gate_persistent_layer_vs_oracle[23-5-24-256-True-rnn_relu]"
run memcheck gru "tests/test_gpu_recurrence.py::test_gru_persistent_layer_vs_oracle[9-40-32-256-True]"
run memcheck lstm64 "tests/test_gpu_recurrence.py::test_single_piece_bf16_recurrence[lstm-19-70-32-256]" "tests/test_gpu_parity.py::test_lstm_persistent_tcgen05_layer_vs_oracle[19-70-32-128-True]"
run memcheck ctc "tests/test_gpu_parity.py" -k "ctc"
run racecheck ctc "tests/test_gpu_parity.py" -k "ctc_tf_known_answer or ctc_host"
run racecheck rnn "tests/test_gpu_recurrence.py::test_one_gate_persistent_layer_vs_oracle[7-1-16-768-True-rnn_tanh]"
run racecheck lstm64 "tests/test_gpu_parity.py::test_lstm_persistent_tcgen05_layer_vs_oracle[19-70-32-128-True]"
