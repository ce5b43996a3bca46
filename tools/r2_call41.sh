#!/bin/bash
timeout 300 python tools/r2_diag_conv.py 2>&1 | tail -20 | cut -c1-600
echo "=== CTCASR_CONV_IMPLICIT=0"
CTCASR_CONV_IMPLICIT=0 timeout 300 python tools/r2_diag_conv.py 2>&1 | head -3 | cut -c1-600
