import sys,json,re
for l in sys.stdin:
    if not l.startswith('{'): print(l.strip()[:150]); continue
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    print('split',d.get('swd_split_waves'), 'S',d.get('swd_searches_per_warp'), d.get('swd_group_searches_per_warp'), 'conc', d.get('concurrent'), 'total', d['total_ms'], d['kernels'], 'rounds', d.get('rounds'), d.get('same_as_first'))
