"""N > 1 host logic on CPU: world_size-2 gloo run of the shard / gather plumbing that bench.py and
multi-GPU callers use.  The per-shard "compute" here is the oracle (CPU); on the GPU box the same
plumbing carries the CUDA results (tests/test_gpu_parity.py covers those)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from cuda_matrix_inversion_b200.sharding import gather_shards, shard_bounds


def test_shard_bounds_cover_the_batch_exactly():
    for batch in (0, 1, 7, 100, 200000, 1 << 20):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(batch, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b
            assert max(b - a for a, b in spans) == -(-batch // world) or batch == 0
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _worker(rank, world, port, batch, n, tmp):
    import oracle as orc
    from tests.util import gp_batch
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    g = gp_batch(n, batch, np.float64, seed=5)                     # same inputs on every rank
    lo, hi = shard_bounds(batch, world, rank)
    means, _ = orc.gp_mean(n, g["a"][lo:hi].reshape(-1), orc.to_colmajor(g["b"][lo:hi]), g["c"][lo:hi].reshape(-1),
                           g["d"][lo:hi].reshape(-1))
    full = gather_shards(torch.from_numpy(means), batch)
    torch.distributed.barrier()
    if rank == 0:
        np.save(os.path.join(tmp, "gathered.npy"), full.numpy())
    torch.distributed.destroy_process_group()


def test_two_rank_gloo_gather_matches_single_rank(tmp_path):
    import oracle as orc
    from tests.util import gp_batch
    batch, n, world = 37, 8, 2                                       # ragged: 19 + 18
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, batch, n, str(tmp_path)), nprocs=world, join=True)
    g = gp_batch(n, batch, np.float64, seed=5)
    want, _ = orc.gp_mean(n, g["a"].reshape(-1), orc.to_colmajor(g["b"]), g["c"].reshape(-1), g["d"].reshape(-1))
    got = np.load(os.path.join(str(tmp_path), "gathered.npy"))
    np.testing.assert_array_equal(got, want)                          # sharding must not change arithmetic
