#!/bin/bash
# usage: dbg_matrix.sh  -- runs dbg_decode.py over a matrix, one process each
run() { echo -n "B=$1 HQ=$2 HKV=$3 S=$4 SPLITS=$5 FUSED=$6: "; B=$1 HQ=$2 HKV=$3 S=$4 OMX_DECODE_SPLITS=$5 FUSED=$6 CUDA_LAUNCH_BLOCKING=1 timeout 120 python scripts/dbg_decode.py 2>&1 | grep -E "max err|fused ok|Error|error" | tr '\n' ' '; echo; }
run 1 4 1 960 1 0
run 1 4 1 8191 1 0
run 1 4 1 8191 9 0
run 4 32 8 8191 0 0
run 4 32 8 8191 0 1
run 4 32 8 8191 1 1
run 1 32 8 8191 0 1
run 4 16 4 2000 0 1
