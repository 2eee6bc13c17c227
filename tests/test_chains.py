"""Chain sharding + posterior pooling (the N > 1 path) on CPU: world_size 2, gloo backend.

BayHunter's chains never interact while sampling (src/mcmcOptimizer.py:208-216); the only
collective of the GPU workflow is the end-of-run all-gather of the fixed-shape posterior blocks that
Plotting.save_final_distribution pools (src/Plotting.py:161-258)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bayhunter_b200 import chains


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, nchains, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = chains.shard_bounds(nchains, rank, world)
        C, S, L, T = hi - lo, 3, 4, 2
        blk = chains.PosteriorBlock(C, S, L, T)
        rng = np.random.default_rng(chains.chain_seed(100, rank))
        for s in range(S):
            rows = torch.from_numpy(rng.uniform(1, 5, (C, L, 4)))
            nlay = torch.from_numpy(rng.integers(2, L + 1, C).astype(np.int32))
            logL = torch.full((C,), float(1000 * rank + s), dtype=torch.float64) + torch.arange(lo, hi, dtype=torch.float64)
            mis = torch.from_numpy(rng.uniform(0, 1, (C, T + 1)))
            noise = torch.from_numpy(rng.uniform(0, 1, (C, 2 * T)))
            blk.record(s, rows, nlay, logL, mis, noise)
        pooled = chains.pool_posterior(blk)
        out_q.put((rank, lo, hi, {k: v.numpy().copy() for k, v in pooled.items()},
                   {k: v.numpy().copy() for k, v in blk.tensors().items()}))
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_all_chains():
    for n in (1, 7, 8, 65536, 21):
        for w in (1, 2, 4, 8):
            b = [chains.shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(120)
def test_pool_posterior_world_size_2_gloo():
    world, nchains = 2, 8
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nchains, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=100) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    # every rank holds the same pooled arrays = the per-rank blocks concatenated in rank order
    for k in chains.PosteriorBlock.FIELDS:
        ref = np.concatenate([r[4][k] for r in res], axis=0)
        for r in res:
            assert r[3][k].shape[0] == nchains
            assert np.array_equal(r[3][k], ref, equal_nan=True), k
    # NaN padding of unused layer slots, float32 storage (src/mcmcOptimizer.py:85-125)
    assert res[0][3]["models"].dtype == np.float32
    assert np.isnan(res[0][3]["models"]).any()


def test_outlier_chains_median_rule():
    likes = torch.tensor([[100.0, 101, 99, float("nan")], [100.5, 100, 101, 100], [50.0, 51, 49, 50]])
    idx, dev = chains.outlier_chains(likes, dev=0.05)
    assert idx.tolist() == [2] and float(dev[0]) > 0.4


def test_record_nuclei_round_trips_through_the_model_adapter():
    """Rows stored with record_nuclei are the reference's parametrisation: Model.get_vp_vs_h gives the layering back."""
    from bayhunter_b200 import Models
    rng = np.random.default_rng(2)
    C, maxl, T = 5, 6, 2
    blk = chains.PosteriorBlock(C, 2, maxl, T)
    models = np.full((C, 2 * maxl), 0.0)
    ks = rng.integers(1, maxl + 1, C)
    for c, k in enumerate(ks):
        models[c, :k] = np.sort(rng.uniform(2, 5, k))
        models[c, maxl:maxl + k] = np.sort(rng.uniform(0, 60, k))
    vpvs = rng.uniform(1.5, 2.0, C)
    blk.record_nuclei(1, torch.from_numpy(models), torch.from_numpy(ks), torch.from_numpy(vpvs),
                      torch.zeros(C, dtype=torch.float64), torch.zeros((C, T + 1), dtype=torch.float64),
                      torch.zeros((C, 2 * T), dtype=torch.float64))
    for c, k in enumerate(ks):
        row = blk.models[c, 1].numpy().astype(np.float64)
        stored = np.concatenate((row[:maxl][:k], row[maxl:][:k]))
        want = np.concatenate((models[c, :k], models[c, maxl:maxl + k])).astype(np.float32).astype(np.float64)
        assert np.array_equal(stored, want)
        assert np.isnan(row[k:maxl]).all() and np.isnan(row[maxl + k:]).all()
        vp, vs, h = Models.Model.get_vp_vs_h(stored, float(blk.vpvs[c, 1]))
        vp2, vs2, h2 = Models.Model.get_vp_vs_h(want, float(np.float32(vpvs[c])))
        assert np.array_equal(h, h2) and np.array_equal(vs, vs2) and np.array_equal(vp, vp2)
