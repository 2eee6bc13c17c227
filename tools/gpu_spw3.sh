#!/bin/bash
ARGS=""
for q in "0 0 0 0" "14 16 26 28" "14 14 28 32" "12 16 28 32" "16 16 21 26" "16 16 22 25" "16 14 24 28" "15 16 24 26" "0 0 0 0"; do
  set -- $q
  ARGS="$ARGS swd_spw_rg=$1,swd_spw_lg=$2,swd_spw_rp=$3,swd_spw_lp=$4,concurrent=1"
done
python tools/quick_bench.py joint5 8192 $ARGS 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if not l.startswith('{'): continue
    d=json.loads(l); print('rg',d['swd_spw_rg'],'lg',d['swd_spw_lg'],'rp', d['swd_spw_rp'], 'lp', d['swd_spw_lp'], 'total', d['total_ms'], 'swd', round(d['kernels']['swd'],2), 'rounds', d['rounds'][1:8:2], d['same_as_first'])
"
