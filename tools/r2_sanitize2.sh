#!/bin/bash
# compute-sanitizer over the kernels added after the first sanitizer run: CTA-pair GEMM, CTC warp kernel with shared-memory
# class sums, im2col to bf16 pieces, dropout, beam-search replay
out=gpurun_out; mkdir -p $out
CS=/usr/local/cuda/bin/compute-sanitizer
run() { tool=$1; shift; name=$1; shift; timeout 900 $CS --tool $tool --print-limit 5 python -m pytest "$@" -x -q > $out/r2san2_${tool}_${name}.log 2>&1; echo "$tool $name: exit $? :: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' $out/r2san2_${tool}_${name}.log | tr '\n' ' ')"; }
run memcheck gemmpair "tests/test_gpu_gemm_pair.py::test_pair_and_single_cta_kernels_agree_on_shared_columns"
run memcheck ctc "tests/test_gpu_parity.py" -k "ctc_ragged or ctc_tf_known or ctc_error"
run memcheck conv "tests/test_gpu_conv.py" -k "ds2_whole_path"
run memcheck dropout "tests/test_gpu_parity.py" -k "dropout_op or rnn_and_dense_dropout"
run memcheck beam "tests/test_gpu_beam.py" -k "artifact"
run racecheck ctc "tests/test_gpu_parity.py" -k "ctc_tf_known_answer or ctc_host"
run racecheck beam "tests/test_gpu_beam.py" -k "artifact"
