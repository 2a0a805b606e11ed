"""Cluster-persistent LSTM / GRU / minimalGRU kernels (csrc/pk_cell_cluster.cu, pk_cell_cluster2.cu) against the step-wise kernels (csrc/pk_cell_step.cu)
through the SAME C-ABI entry points (pk_rnn_step_fwd / pk_rnn_step_bwd; PK_LSTM_CLUSTER selects the family), pytest -m gpu.

The step-wise family is pinned to the oracle and the reference fixtures in tests/test_gpu_parity.py; the full-size
fixture (tests/test_gpu_full_parity.py, config 3) runs on whichever family is the default.  Here: every cluster
geometry (1..5 unit tiles per CTA, cluster sizes 3..14, ragged row counts, one / two directions, each activation
path).  Forward: both families feed identical fp16 operands to mma.sync in the same k order -> bit-identical saved
tensors.  Backward: the K-split partial sums travel as scaled fp16 -> agreement to an fp16 ulp of the carry."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.gpu

SHAPES = [(20, 8, 96, 2, "tanh"), (12, 5, 200, 2, "relu"), (7, 3, 24, 1, "tanh"), (9, 16, 330, 1, "sigmoid"),
          (6, 9, 550, 2, "tanh"), (5, 40, 130, 1, "leaky_relu")]


@pytest.fixture(autouse=True)
def _restore_env():
    old = {k: os.environ.get(k) for k in ("PK_LSTM_CLUSTER", "PK_GRU_CLUSTER")}
    yield
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


@pytest.mark.parametrize("T,B,H,ndir,actn", SHAPES)
def test_cluster_lstm_matches_stepwise(T, B, H, ndir, actn):
    import check_lstm_cluster as chk
    import pk_native as pk
    os.environ["PK_LSTM_CLUSTER"] = "1"
    assert pk.rnn_step_is_cluster(pk.CELL_LSTM, H)
    assert pk.rnn_step_launches(pk.CELL_LSTM, T, B, H, ndir, False) == 2
    act = pk.ACT_IDS[actn]
    d = chk.make(T, B, H, ndir, seed=H + T)
    old, _, _ = chk.run_fwd(d, act, False)
    new, _, _ = chk.run_fwd(d, act, True)
    for k in ("Y32", "Y16", "HT", "HT16", "HP16"):
        assert torch.equal(old[k], new[k]), k
    for i in range(5):
        assert torch.equal(old["SV"][i], new["SV"][i]), f"SV{i}"
    g_old, _ = chk.run_bwd(d, act, old, False)
    g_new, _ = chk.run_bwd(d, act, old, True)
    assert torch.isfinite(g_new.float()).all()
    rel = ((g_old.float() - g_new.float()).norm() / g_old.float().norm().clamp_min(1e-30)).item()
    assert rel < 5e-3, rel


@pytest.mark.parametrize("cell", ["gru", "minimalgru"])
@pytest.mark.parametrize("T,B,H,ndir,actn", [(20, 8, 96, 2, "relu"), (12, 5, 200, 2, "tanh"), (7, 3, 24, 1, "relu"),
                                             (6, 9, 550, 2, "relu"), (5, 40, 130, 1, "sigmoid")])
def test_cluster_gru_matches_stepwise(cell, T, B, H, ndir, actn):
    """GRU / minimalGRU (csrc/pk_cell_cluster2.cu: two exchanges per step) against the step-wise family, same check."""
    import check_lstm_cluster as chk
    import pk_native as pk
    os.environ["PK_GRU_CLUSTER"] = "1"
    cid = chk.CELLS[cell][0]
    assert pk.rnn_step_is_cluster(cid, H)
    assert pk.rnn_step_launches(cid, T, B, H, ndir, True) == 2
    act = pk.ACT_IDS[actn]
    d = chk.make(T, B, H, ndir, seed=H + T, cell=cell)
    old, _, _ = chk.run_fwd(d, act, False)
    new, _, _ = chk.run_fwd(d, act, True)
    for k in ("Y32", "Y16", "HT", "HT16", "HP16", "HX16"):
        assert torch.equal(old[k], new[k]), k
    for i in range(len(old["SV"])):
        assert torch.equal(old["SV"][i], new["SV"][i]), f"SV{i}"
    g_old, _ = chk.run_bwd(d, act, old, False)
    g_new, _ = chk.run_bwd(d, act, old, True)
    assert torch.isfinite(g_new.float()).all()
    rel = ((g_old.float() - g_new.float()).norm() / g_old.float().norm().clamp_min(1e-30)).item()
    assert rel < 5e-3, rel


def test_cluster_lstm_declines_what_it_cannot_hold():
    import pk_native as pk
    os.environ["PK_LSTM_CLUSTER"] = "1"
    assert not pk.rnn_step_is_cluster(pk.CELL_LSTM, 600)      # weight slice + buffers exceed 227 KB of shared memory
    assert not pk.rnn_step_is_cluster(pk.CELL_LIGRU, 2048)    # large-H liGRU stays on the step-wise kernels
    os.environ["PK_LSTM_CLUSTER"] = "0"
    os.environ["PK_GRU_CLUSTER"] = "0"
    assert not pk.rnn_step_is_cluster(pk.CELL_LSTM, 550)
    assert not pk.rnn_step_is_cluster(pk.CELL_GRU, 550)
