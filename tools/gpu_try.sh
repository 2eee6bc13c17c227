#!/bin/bash
python tools/quick_bench.py joint5 8192 concurrent=1 rf_gate_pct=30,concurrent=1 rf_gate_pct=50,concurrent=1 rf_gate_pct=70,concurrent=1 \
   rf_gate_pct=0,swd_spw_lp=32,concurrent=1 rf_gate_pct=30,swd_spw_lp=32,concurrent=1 rf_gate_pct=50,swd_spw_lp=32,concurrent=1 rf_gate_pct=70,swd_spw_lp=32,concurrent=1 rf_gate_pct=85,swd_spw_lp=32,concurrent=1 \
   rf_gate_pct=50,swd_spw_lp=32,swd_spw_rp=32,concurrent=1 rf_gate_pct=70,swd_spw_lp=32,swd_spw_rp=32,concurrent=1 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if not l.startswith('{'): continue
    d=json.loads(l); print({k:v for k,v in d.items() if k.startswith('swd_spw') or k.startswith('rf_')}, 'total', d['total_ms'], {k:round(v,2) for k,v in d['kernels'].items()}, d['same_as_first'])
"
