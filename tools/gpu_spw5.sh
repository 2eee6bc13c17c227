#!/bin/bash
# models-per-warp sweep by curve type (Rayleigh group / Love group / phase), joint5 B = 8192: total ms, swd ms, warps, slowest-warp rounds
mkdir -p gpurun_out
ARGS=""
for rg in 12 14 16; do for lg in 12 14 16; do for ph in 23 26 29 32; do
  ARGS="$ARGS swd_spw_rg=$rg,swd_spw_lg=$lg,swd_spw_rp=$ph,swd_spw_lp=$ph"
done; done; done
python tools/quick_bench.py joint5 8192 concurrent=1 $ARGS 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l[0] != chr(123): continue
    d = json.loads(l)
    if 'skipped' in d: print(d); continue
    r = d['rounds']
    B = 8192
    w = sum((B + d.get(k, s) - 1) // d.get(k, s) for k, s in (('swd_spw_rg', 16), ('swd_spw_lg', 16), ('swd_spw_rp', 23), ('swd_spw_lp', 23)))
    print(d.get('swd_spw_rg'), d.get('swd_spw_lg'), d.get('swd_spw_rp'), 'total %.3f swd %.3f warps %d max rounds %s sum %d' % (d['total_ms'], d['kernels']['swd'], w, [r[1], r[3], r[5], r[7]], r[0] + r[2] + r[4] + r[6]))
" | tee gpurun_out/spw5.txt
