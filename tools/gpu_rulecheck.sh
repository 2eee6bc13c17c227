#!/bin/bash
# the engine's layout rule over batch sizes and configurations (default settings)
for cb in "joint5 64" "joint5 256" "joint5 512" "joint5 1024" "joint5 2048" "joint5 2560" "joint5 3072" "joint5 4096" "joint5 6144" "joint5 8192" "swd2 1024" "swd2 2048" "swd2 4096" "swd2 8192" "swd2 16384" "transd3 256" "transd3 1024" "transd3 2048" "transd3 4096"; do
  set -- $cb
  BH_DEBUG=1 timeout 100 python tools/quick_bench.py $1 $2 swd_pool=-1 2>&1 | python -c "
import sys, json
w = ''
for l in sys.stdin:
    if l.startswith('[bh] swd'): w += l.split(':')[1].split(',')[0].strip() + ' '
    try: d = json.loads(l)
    except Exception: continue
    print('$1 $2 total', d.get('total_ms'), 'swd', d['kernels'].get('swd'), d['kernels'].get('swd_pool'), 'evaluated', d.get('evaluated'), '|', w)
"
done
