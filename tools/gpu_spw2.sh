#!/bin/bash
# clean per-curve-type models-per-warp probes (each line sets every override explicitly)
python tools/quick_bench.py joint5 8192 \
  swd_spw_rg=0,swd_spw_rp=0,swd_spw_lg=0,swd_spw_lp=0,concurrent=1 \
  swd_spw_rg=0,swd_spw_rp=0,swd_spw_lg=0,swd_spw_lp=32,concurrent=1 \
  swd_spw_rg=0,swd_spw_rp=32,swd_spw_lg=0,swd_spw_lp=32,concurrent=1 \
  swd_spw_rg=0,swd_spw_rp=0,swd_spw_lg=8,swd_spw_lp=32,concurrent=1 \
  swd_spw_rg=0,swd_spw_rp=0,swd_spw_lg=0,swd_spw_lp=0,concurrent=1 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if not l.startswith('{'): continue
    d=json.loads(l); print({k[8:]:v for k,v in d.items() if k.startswith('swd_spw')}, 'total', d['total_ms'], 'swd', round(d['kernels']['swd'],2), 'rounds', d['rounds'])
"
