"""Chain sharding over GPUs and posterior pooling.

BayHunter's only parallelism is one OS process per chain writing into its own
slice of shared float32 arrays (src/mcmcOptimizer.py:78-128, 219-252); chains
never interact while sampling.  The B200 equivalent: one process per GPU
(torchrun), a contiguous block of chains per rank, no communication during
sampling, and exactly ONE all-gather at the end that pools the fixed-shape
posterior arrays (models, likes, misfits, noise, vpvs -- NaN padded like the
reference's RawArrays) so that every rank holds what
Plotting.save_final_distribution pools (src/Plotting.py:161-258).

Backend is whatever the process group was created with: "nccl" on GPUs (NVLink 5
/ NVSwitch), "gloo" in the CPU tests.
"""
import numpy as np


def shard_bounds(nchains, rank, world_size):
    """Contiguous block [lo, hi) of chains owned by `rank` (sizes differ by <= 1)."""
    base, rem = divmod(int(nchains), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def chain_seed(base_seed, rank):
    """Per-rank RNG stream (reference: per-chain seeds from one RandomState,
    src/mcmcOptimizer.py:136)."""
    return int(base_seed) + int(rank)


class PosteriorBlock(object):
    """Fixed-shape per-rank posterior storage, float32 and NaN padded exactly like the
    reference's shared arrays: models [C,S,2*maxlayers], likes [C,S],
    misfits [C,S,T+1], noise [C,S,2T], vpvs [C,S]."""

    FIELDS = ("models", "likes", "misfits", "noise", "vpvs")

    def __init__(self, nchains, nsamples, maxlayers, ntargets, device="cpu"):
        import torch
        f = dict(dtype=torch.float32, device=device)
        nan = float("nan")
        self.models = torch.full((nchains, nsamples, 2 * maxlayers), nan, **f)
        self.likes = torch.full((nchains, nsamples), nan, **f)
        self.misfits = torch.full((nchains, nsamples, ntargets + 1), nan, **f)
        self.noise = torch.full((nchains, nsamples, 2 * ntargets), nan, **f)
        self.vpvs = torch.full((nchains, nsamples), nan, **f)

    def record_nuclei(self, s, models, k, vpvs, logL, misfits, noise):
        """Store sample s of every chain in the reference's own parametrisation: `models` [C, 2*maxlayers]
        Voronoi nuclei (vs of the k[c] nuclei, then their depths, as bh_sampler_get_state returns them) and
        the chain's vp/vs.  Rows stored this way are interchangeable with the c_models files: fed through
        Model.get_vp_vs_h they give back the evaluated model."""
        import torch
        C = models.shape[0]
        maxl = self.models.shape[2] // 2
        half = models.shape[1] // 2
        mask = torch.arange(half, device=models.device)[None, :] < k[:, None]
        nanv = torch.full((C, half), float("nan"), dtype=models.dtype, device=models.device)
        # chainmodels[n, :model.size] = model: vs then z, contiguous (src/SingleChain.py:501); here fixed halves
        self.models[:, s, :half] = torch.where(mask, models[:, :half], nanv).to(torch.float32)
        self.models[:, s, maxl:maxl + half] = torch.where(mask, models[:, half:], nanv).to(torch.float32)
        self.likes[:, s] = logL.to(torch.float32)
        self.misfits[:, s] = misfits.to(torch.float32)
        self.noise[:, s] = noise.to(torch.float32)
        self.vpvs[:, s] = vpvs.to(torch.float32)

    def record(self, s, rows, nlay, logL, misfits, noise):
        """Store sample s of every chain from the engine's packed rows + outputs (benchmark / diagnostics).
        The depth half holds LAYER-CENTRE depths, not Voronoi nuclei (nuclei are not unique given the
        interfaces), and vpvs is the top row's: such rows describe the evaluated layering but are NOT
        interchangeable with the reference's c_models files -- use record_nuclei for that."""
        import torch
        C, L, _ = rows.shape
        maxl = self.models.shape[2] // 2
        mask = torch.arange(L, device=rows.device)[None, :] < nlay[:, None]
        nanv = torch.full_like(rows[:, :, 0], float("nan"))
        vs = torch.where(mask, rows[:, :, 0], nanv)
        # Voronoi nucleus depth is not unique given the interfaces; store layer-centre depths
        zc = torch.where(mask, rows[:, :, 2] + 0.5 * rows[:, :, 3], nanv)
        self.models[:, s, :L] = vs.to(torch.float32)
        self.models[:, s, maxl:maxl + L] = zc.to(torch.float32)
        self.likes[:, s] = logL.to(torch.float32)
        self.misfits[:, s] = misfits.to(torch.float32)
        self.noise[:, s] = noise.to(torch.float32)
        self.vpvs[:, s] = rows[:, 0, 1].to(torch.float32)

    def tensors(self):
        return {k: getattr(self, k) for k in self.FIELDS}


def pool_posterior(block, group=None):
    """ONE collective: all ranks' blocks concatenated along the chain axis.

    The five arrays are packed into a single flat float32 buffer so that exactly one
    all_gather_into_tensor crosses NVLink; returns {field: tensor [C_total, ...]}.
    Requires equal chain counts per rank (pad with NaN chains otherwise)."""
    import torch
    import torch.distributed as dist
    parts = block.tensors() if isinstance(block, PosteriorBlock) else dict(block)
    names = list(parts)
    flat = torch.cat([parts[k].reshape(parts[k].shape[0], -1) for k in names], dim=1).contiguous()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        gathered = flat
    else:
        gathered = torch.empty((world * flat.shape[0], flat.shape[1]), dtype=flat.dtype, device=flat.device)
        dist.all_gather_into_tensor(gathered, flat, group=group)
    out, o = {}, 0
    for k in names:
        w = int(np.prod(parts[k].shape[1:])) if parts[k].dim() > 1 else 1
        out[k] = gathered[:, o:o + w].reshape((gathered.shape[0],) + tuple(parts[k].shape[1:]))
        o += w
    return out


def outlier_chains(likes, dev=0.05):
    """Outlier-chain detection by median likelihood (src/Plotting.py:113-154).
    likes [C, S] (NaN padded).  Returns (indices, 1 - score) of chains deviating > dev."""
    import torch
    med = torch.nanmedian(likes.to(torch.float64), dim=1).values
    maxlike = med.max()
    scores = med / maxlike if maxlike > 0 else maxlike / med
    bad = (1 - scores) > dev
    idx = torch.nonzero(bad).flatten()
    return idx, (1 - scores)[idx]


def save_final_distribution(datapath, maxmodels=200000, dev=0.05, rstate=None):
    """Pool the main-phase chain files into the final posterior files, like
    PlotFromStorage.save_final_distribution (src/Plotting.py:161-258): chains whose median
    likelihood deviates more than `dev` from the best chain are dropped (get_outliers :113-154,
    `outliers.dat` written), the others contribute maxmodels / nchains models each (random
    subset, order kept), and c_models / c_likes / c_misfits / c_noise / c_vpvs .npy are saved.
    Reads the per-chain files MCMC_Optimizer wrote; returns the pooled arrays."""
    import glob
    import os
    import os.path as op
    if rstate is None:
        rstate = np.random.RandomState(333)          # module-level rstate of the reference's Plotting.py
    names = ['models', 'likes', 'misfits', 'noise', 'vpvs']
    files = {n: sorted(glob.glob(op.join(datapath, 'c???_p2%s.npy' % n))) for n in names}
    nfiles = {len(v) for v in files.values()}
    if len(nfiles) != 1 or 0 in nfiles:
        raise IOError('missing chain files in %s' % datapath)
    cidx = np.array([int(op.basename(f)[1:4]) for f in files['likes']])
    medians = np.array([np.median(np.load(f)) for f in files['likes']])
    maxlike = np.max(medians)
    scores = medians / maxlike if maxlike > 0 else maxlike / medians
    bad = (1 - scores) > dev
    outliers = cidx[bad]
    outlierfile = op.join(datapath, 'outliers.dat')
    if op.exists(outlierfile):
        os.remove(outlierfile)
    if outliers.size:
        with open(outlierfile, 'w') as f:
            f.write('# Outlier chainindices with %.3f deviation condition\n' % dev)
            for c, sc in zip(outliers, (1 - scores)[bad]):
                f.write('%d\t%.3f\n' % (c, sc))
    nchains = int(cidx.size - outliers.size)
    mpc = int(int(maxmodels) / nchains)
    pooled = {n: [] for n in names}
    for i, c in enumerate(cidx):
        if c in outliers:
            continue
        n = len(np.load(files['likes'][i]))
        index = np.arange(n).astype(int)
        if n > mpc:
            index = rstate.choice(index, mpc, replace=False)
            index.sort()
        for name in names:
            pooled[name].append(np.load(files[name][i])[index])
    out = {}
    for name in names:
        out[name] = np.concatenate(pooled[name], axis=0)
    keep = ~np.isnan(out['likes'])
    for name in names:
        out[name] = out[name][keep]
        np.save(op.join(datapath, 'c_%s' % name), out[name])
    out['outliers'] = outliers
    return out
