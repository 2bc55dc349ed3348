#!/bin/bash
# calibration of the truncation de-bias factors and taps per flush (per-layer signed errors vs the oracle)
run() { echo "== $*"; env "$@" timeout 300 python tests/debug_tc.py 1 2>&1 | grep -E "A1 h\+l|A2 h\+l|feat"; }
run CIA_L2_DEBIAS=0 CIA_L3_DEBIAS=0
run CIA_L2_DEBIAS=1 CIA_L3_DEBIAS=0
run CIA_L2_DEBIAS=0 CIA_L3_DEBIAS=1
run CIA_L2_DEBIAS=0 CIA_L3_DEBIAS=0 CIA_L3_TAPS_PER_FLUSH=3
run CIA_L2_DEBIAS=0 CIA_L3_DEBIAS=2 CIA_L3_TAPS_PER_FLUSH=3
