import sys,json,re
for l in sys.stdin:
    if not l.startswith('{'): print(l.strip()[:150]); continue
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    print({k:v for k,v in d.items() if k.startswith("swd_")}, 'conc', d.get('concurrent'), 'total', d['total_ms'], d['kernels'], 'rounds', d.get('rounds'), d.get('same_as_first'))
