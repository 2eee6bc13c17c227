#!/bin/bash
# tools/gpu_sanitize.sh: compute-sanitizer memcheck + racecheck over the smoke workload and a few lock-step sampler
# iterations (small batches: the tools slow kernels down ~50x).  Logs -> gpurun_out/sanitize_*.log
mkdir -p gpurun_out
cat > /tmp/san_work.py <<'PY'
import numpy as np, sys, os
sys.path.insert(0, os.getcwd())
import __graft_entry__ as g
g.smoke()
import bayhunter_b200 as bh
from bayhunter_b200 import synthetic, Targets as T_, SingleChain as sc
from oracle import joint_oracle as jo
st3 = synthetic.ST3
h, vs = st3["h"], st3["vs"]; vp = vs * st3["vpvs"]; rho = vp * 0.32 + 0.77
per = np.linspace(1, 40, 12); xrf = -5.0 + 0.2 * np.arange(101)
jt = T_.JointTarget([T_.RayleighDispersionPhase(per, jo.surfdisp(h, vp, vs, rho, "rdispph", per)[1]),
                     T_.LoveDispersionGroup(per, jo.surfdisp(h, vp, vs, rho, "ldispgr", per)[1]),
                     T_.PReceiverFunction(xrf, jo.recfunc(h, vp, vs, rho, xrf)[1])])
priors = dict(vs=(2, 5), z=(0, 60), layers=(1, 6), vpvs=(1.4, 2.1), swdnoise_corr=0., swdnoise_sigma=(1e-5, 0.05),
              rfnoise_corr=0.9, rfnoise_sigma=(1e-5, 0.05))
ens = sc.ChainEnsemble(jt, priors, dict(iter_burnin=4, iter_main=4, thickmin=0.1, acceptance=(40, 45), rcond=1e-5),
                       nchains=48, seed=1, max_accepted=8)
ens.init(); ens.run(3)
print("sampler ok", ens.state()["iiter"][:4])
# the lock-step dispersion kernel and the asynchronous host entry
eng, = [bh.Engine([bh.TargetSpec("rdispgr", per, np.full(12, 3.5)), bh.TargetSpec("ldispph", per, np.full(12, 3.6))], 40, 7)]
rows, nlay = synthetic.draw_batch(40, (3, 7), seed=3)
noise = synthetic.draw_noise(40, ("rdispgr", "ldispph"), seed=4)
a = eng.eval_host(rows, nlay, noise)
eng.set(swd_lockstep=1)
b = eng.eval_host(rows, nlay, noise)
assert np.array_equal(a[0], b[0])
out = (np.empty(40), np.empty((40, 3)), np.empty(40, dtype=np.int32), None)
t = eng.submit_host(rows, nlay, noise, out); eng.wait(t)
assert np.array_equal(out[0], a[0])
print("lockstep + async ok")
# the pool dispersion kernel: group + phase curve per wave type, Rayleigh and Love launches side by side
eng2 = bh.Engine([bh.TargetSpec("rdispgr", per, np.full(12, 3.5)), bh.TargetSpec("rdispph", per, np.full(12, 3.5)),
                  bh.TargetSpec("ldispgr", per, np.full(12, 3.6)), bh.TargetSpec("ldispph", per, np.full(12, 3.6))], 70, 7)
rows, nlay = synthetic.draw_batch(70, (3, 7), seed=5)
noise = synthetic.draw_noise(70, ("rdispgr", "rdispph", "ldispgr", "ldispph"), seed=6)
eng2.set(swd_pool=0)
a = eng2.eval_host(rows, nlay, noise)
for m in (0, 9, 28):
    eng2.set(swd_pool=1, swd_pool_models=m)
    b = eng2.eval_host(rows, nlay, noise)
    assert np.array_equal(a[0], b[0])
print("pool ok")
PY
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_work.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok|sampler ok|lockstep|pool ok" gpurun_out/sanitize_$tool.log | tail -6
done
