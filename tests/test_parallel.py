"""Host-side data-parallel logic on CPU: world_size-2 gloo processes shard a batch of image pairs,
run the (oracle) forward on their shard and gather; the result must equal the single-process batch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pwcnet_b200.parallel import gather_pairs, max_over_ranks, shard_range


def test_shard_range_partitions_the_batch():
    for n in (0, 1, 5, 8, 64):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _worker(rank, world, port, n_pairs, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pwc_oracle as O
    torch.set_num_threads(2)
    W = O.glorot_weights(3, gain=1.3, bias_scale=0.02)
    im0, im1 = O.synthetic_pair(n_pairs, 64, 64, 5, shift=(2, -1))
    lo, hi = shard_range(n_pairs, rank, world)
    ff, _ = O.pwcdcnet_forward(W, im0[lo:hi], im1[lo:hi])
    full = gather_pairs(ff, n_pairs)
    slow = max_over_ranks(float(rank + 1))
    if rank == 0:
        ref, _ = O.pwcdcnet_forward(W, im0, im1)
        np.savez(out_path, full=full.numpy(), ref=ref.numpy(), slow=slow)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_forward_equals_single_process(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "r.npz")
    mp.spawn(_worker, args=(2, port, 3, out), nprocs=2, join=True)   # 3 pairs over 2 ranks: ragged shards
    d = np.load(out)
    assert d["full"].shape == d["ref"].shape == (3, 64, 64, 2)
    np.testing.assert_allclose(d["full"], d["ref"], atol=1e-4)      # same pairs in batch order (oneDNN blocking differs with batch size)
    assert float(d["slow"]) == 2.0
