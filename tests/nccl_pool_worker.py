"""Worker of tests/test_gpu_nccl.py (run under torchrun, one rank per GPU, backend nccl): pools posterior blocks
with rank-dependent contents and checks every VALUE of the pooled arrays on every rank; then shards a chain
ensemble over the ranks and checks that a chain's trajectory does not depend on the rank it ran on."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bayhunter_b200 import chains  # noqa: E402


def expected_block(rank, C, S, L, T):
    """Deterministic contents of rank `rank` (reproducible on every rank without communication)."""
    rng = np.random.default_rng(chains.chain_seed(4242, rank))
    rows = rng.uniform(1, 5, (S, C, L, 4))
    nlay = rng.integers(2, L + 1, (S, C)).astype(np.int32)
    logL = rng.normal(-100, 10, (S, C)) + 1000 * rank
    mis = rng.uniform(0, 1, (S, C, T + 1))
    noise = rng.uniform(0, 1, (S, C, 2 * T))
    return rows, nlay, logL, mis, noise


def fill(blk, data, dev):
    rows, nlay, logL, mis, noise = data
    for s in range(rows.shape[0]):
        blk.record(s, torch.from_numpy(rows[s]).to(dev), torch.from_numpy(nlay[s]).to(dev),
                   torch.from_numpy(logL[s]).to(dev), torch.from_numpy(mis[s]).to(dev), torch.from_numpy(noise[s]).to(dev))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    C, S, L, T = 64, 5, 7, 3
    mine = chains.PosteriorBlock(C, S, L, T, device=dev)
    fill(mine, expected_block(rank, C, S, L, T), dev)
    pooled = chains.pool_posterior(mine)
    torch.cuda.synchronize()
    for r in range(world):
        ref = chains.PosteriorBlock(C, S, L, T, device=dev)
        fill(ref, expected_block(r, C, S, L, T), dev)
        for k, v in ref.tensors().items():
            got = pooled[k][r * C:(r + 1) * C]
            assert got.shape == v.shape, (k, got.shape, v.shape)
            assert torch.equal(torch.isnan(got), torch.isnan(v)), k
            assert torch.equal(torch.nan_to_num(got), torch.nan_to_num(v)), "rank %d: pooled %s of rank %d differs" % (rank, k, r)
    assert pooled["likes"].shape[0] == world * C
    # outlier rule on the pooled likelihoods is the same on every rank
    idx, _ = chains.outlier_chains(pooled["likes"], dev=0.05)
    cnt = torch.tensor([idx.numel()], device=dev)
    lo, hi = cnt.clone(), cnt.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert int(lo) == int(hi)
    dist.barrier()
    if rank == 0:
        print("NCCL_POOL_OK world=%d chains=%d" % (world, world * C), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
