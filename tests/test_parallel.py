"""Host-side data-parallel logic on CPU: world_size-2 gloo processes shard a batch of image pairs,
run the (oracle) forward on their shard and gather; the result must equal the single-process batch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pwcnet_b200.parallel import allreduce_gradients, gather_pairs, max_over_ranks, shard_range


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


def _grad_worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pwc_oracle as O
    torch.set_num_threads(2)
    n_pairs = 2
    W0 = O.glorot_weights(4, gain=1.2, bias_scale=0.02)
    im0, im1 = O.synthetic_pair(n_pairs, 64, 64, 6, shift=(1, -2))
    gt = np.random.default_rng(7).normal(0, 3, (n_pairs, 64, 64, 2)).astype(np.float32)
    lo, hi = shard_range(n_pairs, rank, world)

    def grads(a, b):
        Wt = {k: torch.from_numpy(v.copy()).requires_grad_(True) for k, v in W0.items()}
        total, epe, _, _ = O.training_loss(Wt, im0[a:b], im1[a:b], gt[a:b], gamma=0.0)
        total.backward()
        return torch.cat([Wt[k].grad.reshape(-1) for k in sorted(Wt)]), torch.stack([total.detach(), torch.zeros(()), epe.detach()])

    flat, scalars = grads(lo, hi)                    # this rank's shard: what Trainer.forward_backward leaves in grad_flat
    world_n = allreduce_gradients(flat, scalars)     # the one data-path collective of a training step
    flat /= world_n                                  # Trainer folds 1/world into the Adam kernel's grad_scale
    if rank == 0:
        ref, ref_s = grads(0, n_pairs)               # single process, whole batch: L2loss is a mean over the batch
        np.savez(out_path, got=flat.numpy(), ref=ref.numpy(), s=scalars.numpy(), ref_s=ref_s.numpy(), world=world_n)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gradient_allreduce_equals_whole_batch_gradient(tmp_path):
    """Data-parallel training: each rank back-propagates its shard, the flat gradient is summed over
    ranks and scaled by 1/world -> identical to the single-process gradient of the whole batch (the reference's
    losses are means over the batch dimension), and the logged scalars are averaged."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "g.npz")
    mp.spawn(_grad_worker, args=(2, port, out), nprocs=2, join=True)
    d = np.load(out)
    assert int(d["world"]) == 2
    scale = float(np.abs(d["ref"]).max())
    assert float(np.abs(d["got"] - d["ref"]).max()) < 1e-4 * scale
    np.testing.assert_allclose(d["s"][[0, 2]], d["ref_s"][[0, 2]], rtol=1e-4)


def test_allreduce_gradients_is_a_noop_without_a_process_group():
    g = torch.ones(8)
    assert allreduce_gradients(g, torch.ones(3)) == 1 and float(g.sum()) == 8.0
