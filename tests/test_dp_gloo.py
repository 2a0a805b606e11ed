"""world_size-2 data-parallel plumbing on CPU (gloo): flat parameter / gradient buffers, ONE allreduce
per step, 1/world scaling inside the optimizer, replicas stay identical, and the result equals a single
process that sees both shards (mean loss over equal shards = sum of shard gradients / world).

The compute here is a tiny stand-in loss on the drop-in modules' parameters (the CUDA kernels cannot run in
this container); what is under test is pk_train.FlatTrainer's host-side logic and the sharding algebra of
SURVEY.md 8e, with the oracle's RMSprop rule injected in place of the CUDA optimizer kernel."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _setup_paths():
    for p in (os.path.join(ROOT, "pytorch-kaldi_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)


def _oracle_rmsprop(flat_p, flat_g, flat_v, gscale):
    import pk_oracle as orc
    p, v = orc.rmsprop_step(flat_p.numpy().astype(np.float64), flat_g.numpy().astype(np.float64) * gscale,
                            flat_v.numpy().astype(np.float64), lr=0.0004, alpha=0.95, eps=1e-8)
    flat_p.copy_(torch.from_numpy(p).float())
    flat_v.copy_(torch.from_numpy(v).float())


def _model():
    _setup_paths()
    import neural_networks as pknn
    from structure_cases import CASES
    torch.manual_seed(99)
    cls, opts, inp = CASES["ligru_uni_nobn"]
    return pknn.liGRU(dict(opts), inp)


def _surrogate_loss(net, shard):
    # any differentiable function of ALL trainable tensors and of the data shard will do
    tot = 0.0
    for i, p in enumerate(q for q in net.parameters() if q.dim() == 2):
        tot = tot + (p * shard[i % shard.shape[0]].mean()).pow(2).sum() + (p.sum() * shard.std())
    return tot / shard.numel()


def _worker(rank, world, port, out):
    _setup_paths()
    import pk_train
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    net = _model()
    tr = pk_train.FlatTrainer([net], optimizer_fn=_oracle_rmsprop)
    assert tr.world == world
    data = torch.arange(2 * 6 * 5, dtype=torch.float32).reshape(2, 6, 5) / 7.0
    for step in range(3):
        tr.zero_grad()
        _surrogate_loss(net, data[rank] + step).backward()
        # gradients landed in the flat buffer (parameters' .grad are views of it)
        assert all(p.grad is None or p.grad.data_ptr() >= tr.flat_g.data_ptr() for p in net.parameters())
        tr.step()
    out[rank] = tr.flat_p.clone().numpy()
    dist.destroy_process_group()


def test_two_rank_data_parallel_matches_single_process():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert np.array_equal(out[0], out[1]), "replicas diverged"
    # single process: gradient = mean over the two shards' gradients
    _setup_paths()
    import pk_train
    net = _model()
    tr = pk_train.FlatTrainer([net], optimizer_fn=_oracle_rmsprop)
    data = torch.arange(2 * 6 * 5, dtype=torch.float32).reshape(2, 6, 5) / 7.0
    for step in range(3):
        tr.zero_grad()
        (0.5 * (_surrogate_loss(net, data[0] + step) + _surrogate_loss(net, data[1] + step))).backward()
        tr.step()
    ref = tr.flat_p.numpy()
    assert np.max(np.abs(ref - out[0])) <= 1e-6 * max(1.0, np.max(np.abs(ref)))


def test_flat_trainer_keeps_state_dict_api():
    _setup_paths()
    import pk_train
    net = _model()
    before = {k: v.clone() for k, v in net.state_dict().items()}
    tr = pk_train.FlatTrainer([net], optimizer_fn=_oracle_rmsprop)
    after = net.state_dict()
    assert list(before) == list(after)
    assert all(torch.equal(before[k], after[k]) for k in before)
    assert tr.n == sum(p.numel() for p in net.parameters())
    with pytest.raises(RuntimeError):
        pk_train.FlatTrainer([_model()])  # CPU modules without the test hook are refused
