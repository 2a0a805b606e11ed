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


def _cpu_trainer(modules):
    """FlatTrainer with its two device-specific hooks overridden IN THE TEST (the product class has no test hook):
    CPU modules are accepted and the fused CUDA optimizer kernel is replaced by the oracle's RMSprop rule."""
    import pk_oracle as orc
    import pk_train

    class CpuTrainer(pk_train.FlatTrainer):
        def _require_device(self, p):
            pass

        def _apply_update(self, gscale):
            p, v = orc.rmsprop_step(self.flat_p.numpy().astype(np.float64), self.flat_g.numpy().astype(np.float64) * gscale,
                                    self.flat_v.numpy().astype(np.float64), lr=self.lr, alpha=self.alpha, eps=self.eps)
            self.flat_p.copy_(torch.from_numpy(p).float())
            self.flat_v.copy_(torch.from_numpy(v).float())

    return CpuTrainer(modules)


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
    tr = _cpu_trainer([net])
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
    tr = _cpu_trainer([net])
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
    tr = _cpu_trainer([net])
    after = net.state_dict()
    assert list(before) == list(after)
    assert all(torch.equal(before[k], after[k]) for k in before)
    assert tr.n == sum(p.numel() for p in net.parameters())
    with pytest.raises(RuntimeError):
        pk_train.FlatTrainer([_model()])  # the product class refuses CPU modules
    for kw in (dict(opt="rmsprop", momentum=0.9), dict(opt="rmsprop", centered=True), dict(opt="sgd", nesterov=True),
               dict(opt="sgd", weight_decay=1e-4), dict(opt="adam", amsgrad=True), dict(opt="adagrad")):
        with pytest.raises(NotImplementedError):   # unsupported optimizer options are refused, never ignored
            pk_train.FlatTrainer([_model()], **kw)


def test_optimizer_state_round_trips_in_torch_optim_layout():
    """The reference reloads `optimizer_par` for every chunk (core.py:523-535): the flat optimizer state must export
    to / import from the torch.optim.RMSprop state_dict layout, parameter-indexed per architecture."""
    _setup_paths()
    net = _model()
    ref = _model()
    ref.load_state_dict(net.state_dict())
    tr = _cpu_trainer([net])
    opt = torch.optim.RMSprop(ref.parameters(), lr=0.0004, alpha=0.95, eps=1e-8)
    data = torch.arange(6 * 5, dtype=torch.float32).reshape(6, 5) / 7.0
    for step in range(2):
        tr.zero_grad()
        _surrogate_loss(net, data + step).backward()
        tr.step()
        opt.zero_grad()
        _surrogate_loss(ref, data + step).backward()
        opt.step()
    sd = tr.optimizer_state_dicts()[0]
    rsd = opt.state_dict()
    assert sd["param_groups"][0]["params"] == rsd["param_groups"][0]["params"]
    for j, st in rsd["state"].items():  # every state the reference holds exists here with the same values
        assert float(sd["state"][j]["step"]) == float(st["step"])
        assert torch.allclose(sd["state"][j]["square_avg"], st["square_avg"], rtol=1e-5, atol=1e-12)
    # a fresh trainer resumes from the reference's own optimizer state and then tracks the reference
    net2 = _model()
    net2.load_state_dict(ref.state_dict())
    tr2 = _cpu_trainer([net2])
    tr2.load_optimizer_state_dicts([rsd])
    tr2.zero_grad()
    _surrogate_loss(net2, data + 2).backward()
    tr2.step()
    opt.zero_grad()
    _surrogate_loss(ref, data + 2).backward()
    opt.step()
    for a, b in zip(net2.parameters(), ref.parameters()):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-7)
