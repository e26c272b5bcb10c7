"""CPU, world_size 2, gloo: the training collective -- sequences sharded over ranks, ONE all-reduce
(sum) of the flat gradient, identical clip + RMSProp on every rank.  Per-rank
gradients come from autograd through the torch restatement of the reference graph (the checker);
on GPUs they come from the CUDA backward (tests/test_gpu_training.py pins that to the same autograd)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ntm_tracker_b200.sharding import shard_range
from oracle import ntm_oracle as O
from oracle.ntm_ref_torch import TorchRefNTM

SHAPE = dict(output_dim=2, input_dim=5, mem_size=12, mem_dim=8, controller_hidden_size=10,
             controller_num_layers=1, write_head_size=1, read_head_size=2)
GATHER = [3, 5]


def ref_grads(params, x, targets):
    s = O.NTMShape(**SHAPE)
    ref = TorchRefNTM(s, params, dtype=torch.float64, requires_grad=True)
    _, logits, _ = ref.run(torch.from_numpy(x))
    loss = 0.5 * torch.sum((torch.tanh(logits[:, GATHER]) - torch.from_numpy(targets).double()) ** 2)
    loss.backward()
    return {k: v.grad.float() for k, v in ref.p.items()}


def _numpy_update(self):
    """Test stand-in for NTMTrainer._fused_update (the product's is a CUDA library call): the same
    tf.clip_by_global_norm + RMSProp arithmetic (direct_offset_output.py:606-626) on the flat CPU buffers."""
    g = self._grad
    gn = torch.linalg.vector_norm(g)
    g = g * (self.clip / torch.clamp(gn, min=self.clip))
    self._rms.mul_(self.decay).addcmul_(g, g, value=1.0 - self.decay)
    self._mom.mul_(self.momentum).add_(self.lr * g / torch.sqrt(self._rms + self.eps))
    self._flat.sub_(self._mom)
    self.last_gnorm.copy_(gn.reshape(1))


def make_trainer(params):
    from ntm_tracker_b200 import LoopNTMTracker, NTMTrainer
    kw = {k: v for k, v in SHAPE.items() if k not in ("input_dim", "output_dim")}
    trk = LoopNTMTracker(6, 2, device="cpu", **kw)     # variables on the CPU: only the host-side logic runs here
    trk.cell.load_reference_weights(params)
    trainer = NTMTrainer(trk, learning_rate=1e-2)
    trainer._fused_update = _numpy_update.__get__(trainer)          # flat buffers, all-reduce, views: the product's
    return trainer


def data():
    s = O.NTMShape(**SHAPE)
    params = O.init_params(s, 9, 0.3, random_biases=True)
    rng = np.random.RandomState(2)
    x = rng.standard_normal((6, 6, 5)).astype(np.float32)
    targets = rng.uniform(-0.5, 0.5, (6, 2, 2)).astype(np.float32)
    return params, x, targets


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        params, x, targets = data()
        lo, hi = shard_range(x.shape[0], world, rank)
        trainer = make_trainer(params)
        for _ in range(2):
            cur = {k: v.detach().numpy() for k, v in trainer.cell.variables.items()}
            g = ref_grads(cur, x[lo:hi], targets[lo:hi])            # this rank's sequences only
            trainer.apply_gradients(g)                               # all-reduce + clip + RMSProp
        if rank == 0:
            np.savez(out, **{k.replace("/", "|"): v.numpy() for k, v in trainer.cell.variables.items()})
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_sharded_training_equals_single_process(tmp_path):
    out = str(tmp_path / "vars.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    z = np.load(out)
    params, x, targets = data()
    trainer = make_trainer(params)                                   # not distributed: full batch
    for _ in range(2):
        cur = {k: v.detach().numpy() for k, v in trainer.cell.variables.items()}
        trainer.apply_gradients(ref_grads(cur, x, targets))
    for k, v in trainer.cell.variables.items():
        np.testing.assert_allclose(z[k.replace("/", "|")], v.numpy(), atol=2e-6, err_msg=k)
