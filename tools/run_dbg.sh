#!/bin/bash
export CUDA_LAUNCH_BLOCKING=1
for v in 8 11 10 9; do
  echo "=== c1_etqr_boston dbg $v"
  timeout 120 python tools/dbg_select4.py c1_etqr_boston select4_dbg=$v 2>&1 | tail -3 | cut -c1-200
done
echo "=== unroll 4, dbg 9"
timeout 120 python tools/dbg_select4.py c1_etqr_boston select4_dbg=9 select4_unroll=4 2>&1 | tail -3 | cut -c1-200
