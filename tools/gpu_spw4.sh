#!/bin/bash
# models-per-warp sweep after the leaner layer loops (rg lg rp lp; 0 = rule)
CFG=${1:-joint5}; B=${2:-8192}
ARGS=""
for q in "0 0 0 0" "16 16 16 16" "16 16 20 20" "14 16 23 23" "16 16 23 28" "14 14 23 23" "16 16 24 24" "12 16 20 24" "16 16 26 26" "0 0 0 0"; do
  set -- $q
  ARGS="$ARGS swd_spw_rg=$1,swd_spw_lg=$2,swd_spw_rp=$3,swd_spw_lp=$4,concurrent=1"
done
python tools/quick_bench.py $CFG $B $ARGS 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if not l.startswith('{'): print(l.strip()[:300]); continue
    d=json.loads(l)
    if 'total_ms' not in d: print(l.strip()); continue
    print('rg',d['swd_spw_rg'],'lg',d['swd_spw_lg'],'rp', d['swd_spw_rp'], 'lp', d['swd_spw_lp'], 'total', d['total_ms'], 'swd', round(d['kernels']['swd'],2), 'rounds', d['rounds'][1:8:2], d['same_as_first'])
"
