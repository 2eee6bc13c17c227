#!/bin/bash
# Round record on one B200: parity tests, smoke, bench (both arms), ncu launch list, ncu full captures of the
# dominant kernels.  Everything lands in gpurun_out/<tag>_*.
mkdir -p gpurun_out
T=${1:-rec}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/${T}_smi.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -3 gpurun_out/${T}_pytest.log
python __graft_entry__.py --smoke > gpurun_out/${T}_smoke.log 2>&1; tail -1 gpurun_out/${T}_smoke.log
python bench.py --steps 100 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 3600 gpurun_out/${T}_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2>> gpurun_out/${T}_bench.err; tail -c 1200 gpurun_out/${T}_bench_ref.json
# every launch of the same command with its device time (cold cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --sampler-iters 2 > gpurun_out/${T}_ncu_bench.log 2>&1
# the dominant kernels once, full set.  A full joint5 batch runs the dispersion search as two concurrent launches of
# swd_pool_kernel per evaluation (Rayleigh first, then Love): skip three evaluations, take the fourth one's two launches
# (ncu serialises them: each capture is that kernel alone on the device)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:swd_pool_kernel -s 6 -c 1 -f -o gpurun_out/${T}_swd \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs --sampler-iters 0 > gpurun_out/${T}_ncu_swd.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:swd_pool_kernel -s 7 -c 1 -f -o gpurun_out/${T}_swdl \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs --sampler-iters 0 > gpurun_out/${T}_ncu_swdl.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rf_spectrum -s 3 -c 1 -f -o gpurun_out/${T}_rf \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs --sampler-iters 0 > gpurun_out/${T}_ncu_rf.log 2>&1
ls -la gpurun_out | tail -14
