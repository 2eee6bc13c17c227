"""MCMC_Optimizer with BayHunter's interface (src/mcmcOptimizer.py:31-300), chains on the GPU.

    optimizer = MCMC_Optimizer(targets, initparams=initparams, priors=priors, random_seed=None)
    optimizer.mp_inversion(nthreads=..., baywatch=..., dtsend=...)

Instead of one OS process per chain, all chains of this rank advance in lock step inside
`SingleChain.ChainEnsemble` (device sampler).  What reaches the disk is what the reference
writes: `<savepath>/data/<station>_config.pkl` (utils.save_config, src/utils.py:127-153) and per
chain `c%03d_p{1,2}{models,likes,misfits,noise,vpvs}.npy` (SingleChain.save_finalmodels), so
PlotFromStorage / save_final_distribution work on the output unchanged.

Multi-GPU: one process per GPU (torchrun); chains are split in contiguous blocks
(chains.shard_bounds), every rank writes the files of its own chains, no communication while
sampling.  `nthreads` and `baywatch` are accepted for interface compatibility (BayWatch's live
socket feed is out of scope).
"""
import logging
import os
import os.path as op
import pickle
import time

import numpy as np

from . import SingleChain as _sc
from .chains import shard_bounds

logger = logging.getLogger()


def save_config(targets, configfile, priors=dict(), initparams=dict()):
    """utils.save_config (src/utils.py:127-153)."""
    data = {}
    refs = []
    for target in targets.targets:
        target.get_covariance = None
        refs.append(target.ref)
    data['targets'] = targets.targets
    data['targetrefs'] = refs
    data['priors'] = priors
    data['initparams'] = initparams
    with open(configfile, 'wb') as f:
        pickle.dump(data, f)


class MCMC_Optimizer(object):
    def __init__(self, targets, initparams=dict(), priors=dict(), random_seed=None, rank=None, world_size=None,
                 max_accepted=None):
        self.rstate = np.random.RandomState(random_seed)
        self.priors = dict(_sc.DEFAULT_PRIORS); self.priors.update(priors)
        self.initparams = dict(_sc.DEFAULT_INITPARAMS); self.initparams.update(initparams)
        self.station = self.initparams.get('station')
        self.targets = targets
        if "LOCAL_RANK" in os.environ:                 # one process per GPU
            from . import _lib
            _lib.set_device(int(os.environ["LOCAL_RANK"]))
        self.rank = int(os.environ.get("RANK", 0)) if rank is None else int(rank)
        self.world_size = int(os.environ.get("WORLD_SIZE", 1)) if world_size is None else int(world_size)

        self.savepath = op.join(self.initparams['savepath'], 'data')
        if not op.exists(self.savepath):
            os.makedirs(self.savepath, exist_ok=True)
        if self.rank == 0:
            outfile = op.join(self.savepath, '%s_config.pkl' % self.station)
            save_config(targets, outfile, priors=self.priors, initparams=self.initparams)

        self.nchains = int(self.initparams.get('nchains'))
        self.ntargets = len(targets.targets)
        self.iter_phase1 = int(self.initparams['iter_burnin'])
        self.iter_phase2 = int(self.initparams['iter_main'])
        self.iterations = self.iter_phase1 + self.iter_phase2
        self.maxlayers = int(self.priors['layers'][1]) + 1

        # per-chain seeds exactly like _init_chain (src/mcmcOptimizer.py:130-138)
        self.chain_seeds = np.array([self.rstate.randint(1000) for _ in range(self.nchains)])
        lo, hi = shard_bounds(self.nchains, self.rank, self.world_size)
        self.chain_range = (lo, hi)
        device_seed = int(random_seed) if random_seed is not None else int(np.random.SeedSequence().entropy % (1 << 63))
        self.ensemble = _sc.ChainEnsemble(targets, self.priors, self.initparams, nchains=hi - lo, first_chain=lo,
                                          seed=device_seed, chain_seeds=self.chain_seeds[lo:hi],
                                          max_accepted=max_accepted)
        self.nmodels = self.ensemble.nmodels
        logger.info('> %d chain(s) are initiated ...' % self.nchains)

    def mp_inversion(self, baywatch=False, dtsend=0.5, nthreads=0, chunk=2048):
        """Run every chain of this rank for iter_burnin + iter_main iterations and save the
        reference's per-chain files.  Returns the wall-clock seconds of the sampling."""
        if baywatch:
            logger.info('BayWatch live feed is not provided by the GPU optimizer.')
        t0 = time.time()
        ens = self.ensemble
        ens.init()
        ens.run_all(chunk=chunk)
        runtime = time.time() - t0
        logger.info('> All chains terminated after: %.5f s' % runtime)
        st = ens.state()
        nover = int((st["overflow_count"] > 0).sum())
        if nover > 0:
            # the record of such a chain is complete only up to its first unstored model: its files are written
            # with that iteration as the end of the last stored model's dwell time (weights stay correct)
            logger.warning('%d chains accepted more models than the chain arrays hold (max_accepted = %d, %d models '
                           'not stored): their saved posterior ends at the iteration of the first unstored model'
                           % (nover, self.nmodels, int(st["overflow"][0])))
        maxmodels = float(self.initparams['maxmodels'])
        lo, hi = self.chain_range
        block = 64
        self.saved = {}
        for c0 in range(0, hi - lo, block):
            n = min(block, hi - lo - c0)
            arr = ens.chain_arrays(c0, n)
            for j in range(n):
                one = {k: v[j] for k, v in arr.items()}
                try:
                    final_iter = self.iter_phase2
                    if st["overflow_count"][c0 + j] > 0:
                        final_iter = min(final_iter, int(st["overflow_iter"][c0 + j]))
                    self.saved[lo + c0 + j] = _sc.save_chain_files(
                        one, int(st["nstored"][c0 + j]), lo + c0 + j, self.savepath, maxmodels, final_iter)
                except ValueError:
                    logger.info('No main phase models accepted.')
        logger.info('### time for inversion: %.2f s' % (time.time() - t0))
        self.state = st
        return runtime
