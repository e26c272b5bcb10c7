"""CPU, world_size 2, gloo: the N>1 host logic -- contiguous sharding of the sequence
batch with no data-path collective, max-over-ranks timing, gather back into batch
order.  The per-rank compute here is the NumPy oracle (the checker); on GPUs each
rank runs the CUDA path on its shard instead (bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ntm_tracker_b200.sharding import gather_batch, max_over_ranks, shard_range
from oracle import ntm_oracle as O


def test_shard_ranges_partition_the_batch():
    for B in (1, 2, 7, 64, 4096, 4097):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(B, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, B, T, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        s = O.NTMShape(output_dim=3, input_dim=6, mem_size=16, mem_dim=8, controller_hidden_size=10,
                       controller_num_layers=1, write_head_size=1, read_head_size=2)
        params = O.init_params(s, 5, 0.05)                   # replicated weights
        x = np.random.RandomState(6).standard_normal((B, T, 6)).astype(np.float32)
        lo, hi = shard_range(B, world, rank)
        _, logits, st = O.run_sequence(params, s, x[lo:hi])   # this rank's sequences only
        full = gather_batch(torch.from_numpy(logits), B)
        full_M = gather_batch(torch.from_numpy(st["M"]), B)
        t = max_over_ranks(1.0 + rank)                        # timing reduction: max over ranks
        dist.barrier()
        if rank == 0:
            np.savez(out_path, logits=full.numpy(), M=full_M.numpy(), t=t)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [5, 8])
def test_two_rank_sharded_run_equals_single_run(tmp_path, B):
    T, world = 4, 2
    out = str(tmp_path / "out.npz")
    mp.spawn(_worker, args=(world, _free_port(), B, T, out), nprocs=world, join=True)
    z = np.load(out)
    s = O.NTMShape(output_dim=3, input_dim=6, mem_size=16, mem_dim=8, controller_hidden_size=10,
                   controller_num_layers=1, write_head_size=1, read_head_size=2)
    params = O.init_params(s, 5, 0.05)
    x = np.random.RandomState(6).standard_normal((B, T, 6)).astype(np.float32)
    _, logits, st = O.run_sequence(params, s, x)
    np.testing.assert_allclose(z["logits"], logits, atol=1e-12)
    np.testing.assert_allclose(z["M"], st["M"], atol=1e-12)
    assert float(z["t"]) == 2.0
