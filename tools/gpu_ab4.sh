#!/bin/bash
# several library builds over a few configurations: tools/gpu_ab4.sh "<cfg B>;<cfg B>..." lib1 lib2 ... [-- settings...]
IFS=';' read -ra CFGS <<< "$1"; shift
LIBS=(); while [ $# -gt 0 ] && [ "$1" != "--" ]; do LIBS+=("$1"); shift; done; shift
SET=("$@"); [ ${#SET[@]} -eq 0 ] && SET=(swd_pool=-1)
for cb in "${CFGS[@]}"; do
  set -- $cb
  for lib in "${LIBS[@]}"; do
    BH_B200_LIB=bayhunter_b200/variants/libbh_$lib.so timeout 120 python tools/quick_bench.py $1 $2 "${SET[@]}" 2>&1 | grep -v "^#" | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()[:200]); continue
    k = d.get('kernels', {})
    print('$1 $2 $lib', {a: d[a] for a in d if a.startswith('swd_')}, 'total', d.get('total_ms'), 'swd', k.get('swd'), k.get('swd_pool'), 'evaluated', d.get('evaluated'), 'same', d.get('same_as_first'))
"
  done
done
