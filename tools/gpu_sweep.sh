#!/bin/bash
# sweep of searches-per-warp settings for the three benchmark configurations (current in-tree library)
mkdir -p gpurun_out
run() { python tools/quick_bench.py "$@" 2>&1 | grep cfg | python tools/fmt_ab.py; }
echo "== joint5 8192"
S=""; for s in "32 16" "16 16" "16 8" "8 8"; do set -- $s; S="$S swd_searches_per_warp=$1,swd_group_searches_per_warp=$2,concurrent=1"; done
run joint5 8192 $S
echo "== swd2 4096"
S=""; for s in "32 16" "16 16" "16 8" "8 8" "8 4" "4 4" "4 2" "2 2"; do set -- $s; S="$S swd_searches_per_warp=$1,swd_group_searches_per_warp=$2,concurrent=1"; done
run swd2 4096 $S
echo "== transd3 4096"
S=""; for s in "16 16" "16 8" "8 8" "8 4" "4 4" "4 2"; do set -- $s; S="$S swd_searches_per_warp=$1,swd_group_searches_per_warp=$2,concurrent=1"; done
run transd3 4096 $S
