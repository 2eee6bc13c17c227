"""bayhunter_b200 -- B200-native forward-model + likelihood engine for BayHunter's
hot path (surface-wave dispersion, receiver functions, Gaussian log-likelihood).

The arithmetic lives in libbayhunter_b200.so (hand-written sm_100a CUDA behind a
C ABI, include/bayhunter_b200.h).  This package is the Python host side: the
ctypes binding, drop-in `SurfDisp` / `RFminiModRF` plugins, a `Targets` module
with BayHunter's interface plus `JointTarget.evaluate_batch`, the model packer,
the device sampler's host side (`SingleChain.ChainEnsemble`, `MCMC_Optimizer`),
`SynthObs`, and the multi-GPU chain sharding / posterior pooling helpers.
"""
from . import _lib
from .engine import Engine, TargetSpec, gauss_corr_inverse
from .Models import Model, pack_layers, pack_models
from .rfmini_modrf import RFminiModRF
from .surf96_modsw import SurfDisp
from . import Targets
from . import SingleChain
from .SingleChain import ChainEnsemble
from .mcmcOptimizer import MCMC_Optimizer
from .SynthObs import SynthObs

__all__ = ["Engine", "TargetSpec", "gauss_corr_inverse", "Model", "pack_layers", "pack_models",
           "RFminiModRF", "SurfDisp", "Targets", "SingleChain", "ChainEnsemble",
           "MCMC_Optimizer", "SynthObs", "_lib"]
