#!/bin/bash
set -u
O=gpurun_out
L=$O/r2_dbg.log
: > $L
timeout 600 python tests/diag_dw_gemm.py >> $L 2>&1
tail -5 $L
