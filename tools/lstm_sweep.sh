#!/bin/bash
# sweep of the persistent-LSTM tuning knobs (tensor-memory-resident k-blocks, L2 keep share, direction stagger);
# every configuration is also checked for bit-exact reproducibility by tools/lstm_check.py
for cfg in "$@"; do
  set -- $(echo $cfg | tr ',' ' ')
  echo -n "KRES=$1 KEEP_MB=$2 STAGGER_NS=$3: "
  CTCASR_LSTM_KRES=$1 CTCASR_LSTM_L2_KEEP_MB=$2 CTCASR_LSTM_STAGGER_NS=$3 timeout 200 python tools/lstm_check.py ${T:-600} ${R:-4} 2>&1 | tail -1 | cut -c1-200
done
