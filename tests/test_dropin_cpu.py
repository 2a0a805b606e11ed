"""CPU tests of the drop-in boundary: the classes of pytorch-kaldi_b200/neural_networks.py must
register exactly what the reference constructors register (state_dict keys / shapes / order / values
under a fixed seed, parameter order, generator consumption, out_dim), must refuse CPU tensors loudly,
and the C-ABI library must load and export every symbol include/pk_b200.h declares."""
import ctypes
import hashlib
import json
import os
import sys
import re

import numpy as np
import pytest
import torch

import neural_networks as pknn
import pk_native
from structure_cases import CASES

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "structure.json")))


def digest(t):
    return hashlib.sha256(t.detach().contiguous().numpy().tobytes()).hexdigest()[:16]


@pytest.mark.parametrize("name", sorted(CASES))
def test_constructor_matches_reference(name):
    cls, opts, inp_dim = CASES[name]
    ref = GOLD[name]
    torch.manual_seed(1234)
    m = getattr(pknn, cls)(dict(opts), inp_dim)
    assert int(m.out_dim) == ref["out_dim"]
    got = [[k, list(v.shape), str(v.dtype), digest(v)] for k, v in m.state_dict().items()]
    assert [g[:3] for g in got] == [r[:3] for r in ref["keys"]], "state_dict keys/shapes/order differ"
    bad = [g[0] for g, r in zip(got, ref["keys"]) if g[3] != r[3]]
    assert not bad, f"initial values differ from the reference for {bad[:5]}"
    assert [[k, list(p.shape)] for k, p in m.named_parameters()] == ref["params"]
    # the constructors consumed the CPU generator exactly like the reference's
    assert float(torch.rand(1).item()) == ref["next_rand"]


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference checkout not present")
def test_live_reference_state_dict_roundtrip():
    """A checkpoint written by the reference loads into the drop-in and vice versa (core.py:531, :715)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("ref_neural_networks", "/root/reference/neural_networks.py")
    ref_nn = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_nn)
    cls, opts, inp_dim = CASES["ligru_timit"]
    torch.manual_seed(7)
    a = ref_nn.liGRU(dict(opts), inp_dim)
    torch.manual_seed(8)
    b = pknn.liGRU(dict(opts), inp_dim)
    b.load_state_dict(a.state_dict())
    for (ka, va), (kb, vb) in zip(a.state_dict().items(), b.state_dict().items()):
        assert ka == kb and torch.equal(va, vb)
    opt_a = torch.optim.RMSprop(a.parameters(), lr=0.0004, alpha=0.95, eps=1e-8)
    opt_b = torch.optim.RMSprop(b.parameters(), lr=0.0004, alpha=0.95, eps=1e-8)
    opt_b.load_state_dict(opt_a.state_dict())  # index-ordered param groups line up


def test_cpu_tensors_are_refused_loudly():
    cls, opts, inp_dim = CASES["ligru_uni_nobn"]
    m = pknn.liGRU(dict(opts), inp_dim)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(4, 2, inp_dim))
    h = pknn.MLP(dict(CASES["mlp_head"][1]), 110)
    with pytest.raises(RuntimeError, match="CUDA"):
        h(torch.zeros(3, 110))


CUDNN_OPTS = dict(hidden_size="24", num_layers="2", bias="True", batch_first="True", dropout="0.2", bidirectional="True",
                  nonlinearity="tanh", use_cuda="False", to_do="train")


def test_cudnn_layout_classes_construct_like_the_reference():
    """LSTM_cudnn / GRU_cudnn / RNN_cudnn (reference :153-297): same sub-module, state_dict keys / shapes / values and
    generator consumption as the reference constructors (cfg/TIMIT_baselines/TIMIT_LSTM_fmllr_cudnn.cfg drops in);
    CPU tensors are refused, nn.GRU's different gate algebra is refused instead of silently mis-computed."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    ref = None
    if os.path.exists(os.path.join(ref_dir, "neural_networks.py")):
        import importlib.util
        spec = importlib.util.spec_from_file_location("ref_nn_for_cudnn", os.path.join(ref_dir, "neural_networks.py"))
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
    for name, attr in (("LSTM_cudnn", "lstm"), ("GRU_cudnn", "gru"), ("RNN_cudnn", "rnn")):
        torch.manual_seed(5)
        m = getattr(pknn, name)(dict(CUDNN_OPTS), 10)
        after = torch.rand(1).item()
        assert m.out_dim == 48
        keys = list(m.state_dict().keys())
        assert keys[0] == f"{attr}.0.weight_ih_l0" and f"{attr}.0.weight_hh_l1_reverse" in keys
        if ref is not None:
            torch.manual_seed(5)
            r = getattr(ref, name)(dict(CUDNN_OPTS), 10)
            assert torch.rand(1).item() == after, "constructor consumed the generator differently"
            assert list(r.state_dict().keys()) == keys
            for k, v in r.state_dict().items():
                assert torch.equal(v, m.state_dict()[k]), (name, k)
    with pytest.raises(RuntimeError, match="CUDA"):
        pknn.LSTM_cudnn(dict(CUDNN_OPTS), 10)(torch.zeros(5, 2, 10))
    with pytest.raises(NotImplementedError):
        pknn.GRU_cudnn(dict(CUDNN_OPTS), 10)(torch.zeros(5, 2, 10))


def test_abi_library_loads_and_exports_every_declared_symbol():
    """No compute calls here (no GPU in this container): only dlopen + symbol resolution."""
    header = open(os.path.join(ROOT, "include", "pk_b200.h")).read()
    declared = set(re.findall(r"\b(pk_[a-z0-9_]+)\s*\(", header))
    assert {"pk_gemm_tn", "pk_rnn_layer_fwd", "pk_rnn_layer_bwd", "pk_logsoftmax_nll"} <= declared
    assert os.path.exists(pk_native.LIB_PATH), "build the library first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(pk_native.LIB_PATH)
    for sym in sorted(declared):
        assert hasattr(lib, sym), f"libpk_b200.so does not export {sym}"
    # every symbol the Python binding declares must be in the header too
    assert set(pk_native.SIGNATURES) | {"pk_last_error", "pk_version"} == declared
    # the launch-count table behind bench.py's gpu_launches covers every compute entry point
    assert all(isinstance(v, int) for v in pk_native.KERNELS_PER_CALL.values())
    assert set(pk_native.SIGNATURES) - set(pk_native.KERNELS_PER_CALL) <= {"pk_rnn_step_workspace_bytes", "pk_rnn_step_launches", "pk_rnn_step_is_cluster"}
    L = pk_native.lib()
    assert L.pk_version() >= 2
    assert L.pk_last_error() is not None


def test_batch_descriptors_follow_the_reference_bookkeeping():
    """pk_train.batch_descriptors (host logic of the device-side batch assembly): same sentence ranges, same
    random left padding (one `randint` per sentence, in order) as core.py:581-598."""
    import random
    import numpy as np
    import golden_util as gu
    import pk_train
    d = gu.load("input_cw")
    m = d["meta"]
    rng = random.Random(m["seed"])
    snt, beg = 0, 0
    for i in range(m["n_snt"] // m["batch"]):
        desc, max_len, snt, beg = pk_train.batch_descriptors(d["data_end_index"], snt, beg, m["batch"], rng)
        ref = d[f"inp{i}"]
        assert max_len == ref.shape[0] and tuple(desc.shape) == (3, m["batch"])
        ds = d["data_set"].astype(np.float32)
        for k in range(m["batch"]):
            b, L, z = (int(v) for v in desc[:, k])
            assert np.array_equal(ref[z:z + L, k], ds[b:b + L])
            assert not ref[:z, k].any() and not ref[z + L:, k].any()


def test_kaldi_matrix_writer_and_counts_reader(tmp_path):
    """Host half of the posterior writer: pk_train.write_kaldi_matrix / load_counts against the reference-written archive
    (the device half, the prior subtraction kernel, is covered by the GPU test)."""
    import io
    import numpy as np
    import golden_util as gu
    import pk_train
    d = gu.load("post_ark")
    cf = tmp_path / "counts"
    cf.write_text("[ " + " ".join(str(int(c)) for c in d["counts"]) + " ]\n")
    counts = pk_train.load_counts(str(cf))
    assert counts.dtype == np.float32 and np.array_equal(counts, d["counts"])
    out = d["logp"] - np.log(counts / np.sum(counts))
    buf = io.BytesIO()
    pk_train.write_kaldi_matrix(buf, "utt_0001", out)
    pk_train.write_kaldi_matrix(buf, "utt_0002", d["logp"][:3])
    assert buf.getvalue() == d["ark"].tobytes()


def test_kaldi_archive_readers_host_side():
    """pk_train.read_mat_ark (FM / DM entries) and read_vec_int_ark against what the reference's readers return for
    the same bytes; a compressed entry without a device is refused loudly (it is decoded on the GPU)."""
    import io
    import numpy as np
    import pytest
    import golden_util as gu
    import pk_train
    d = gu.load("ark_read")
    it = pk_train.read_mat_ark(io.BytesIO(d["feats"].tobytes()))
    k, m = next(it)
    assert k == "utt_fm" and m.dtype == np.float32 and np.array_equal(m, d["mat.utt_fm"])
    k, m = next(it)
    assert k == "utt_dm" and m.dtype == np.float64 and np.array_equal(m, d["mat.utt_dm"])
    with pytest.raises(RuntimeError):
        next(it)
    alis = dict(pk_train.read_vec_int_ark(io.BytesIO(d["alis"].tobytes())))
    assert list(alis) == ["utt_a", "utt_b"]
    for k, v in alis.items():
        assert v.dtype == np.int32 and np.array_equal(v, d["ali." + k])


def test_kaldi_scp_reader(tmp_path):
    """pk_train.read_mat_scp: `key path:offset` lines (data_io.py:1039-1059 / :696-716) into the same archive bytes."""
    import numpy as np
    import golden_util as gu
    import pk_train
    d = gu.load("ark_read")
    blob = d["feats"].tobytes()
    ark = tmp_path / "feats.ark"
    ark.write_bytes(blob)
    lines = []
    for key in ("utt_fm", "utt_dm"):
        off = blob.index(key.encode() + b" ") + len(key) + 1
        lines.append(f"{key} {ark}:{off}")
    scp = tmp_path / "feats.scp"
    scp.write_text("\n".join(reversed(lines)) + "\n")      # random access: order differs from the archive
    got = dict(pk_train.read_mat_scp(str(scp)))
    assert list(got) == ["utt_dm", "utt_fm"]
    for k, m in got.items():
        assert np.array_equal(m, d["mat." + k])


def test_c_abi_error_convention_without_gpu():
    """SURVEY 8b error convention: every entry point returns an int status (0 = ok), never aborts, and leaves a
    thread-local message for pk_last_error().  Argument validation happens before any device work, so it can be
    exercised in this GPU-less container."""
    import ctypes
    L = pk_native.lib()
    rc = L.pk_gemm_tn(0, 0, 8, 8, None, 8, 0, 0, None, 8, 0, 0, None, 8, None, 0, None, 1.0, None, 0, 1, None, None)
    assert rc != 0 and b"null operand" in L.pk_last_error()
    rc = L.pk_rnn_layer_fwd(7, 4, 2, 8, 1, 0, None, 8, None, None, None, None, 1.0, None, 8, None, 8, None, None, None, None,
                            None, 8, None)
    assert rc != 0 and b"not implemented" in L.pk_last_error()
    rc = L.pk_rnn_step_fwd(4, 4, 2, 8, 3, 0, None, 8, None, None, None, None, 1.0, None, 8, None, 8, None, None, None, None,
                           None, None, None, None, None, 8, None, 0, None)
    assert rc != 0 and L.pk_last_error() != b""
    rc = L.pk_sinc_filters_fwd(None, None, 4, 8, 16000.0, 50.0, 50.0, None, None)
    assert rc != 0 and b"null pointer" in L.pk_last_error()
    rc = L.pk_adam_step(None, None, None, None, 10, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1, 1.0, None)
    assert rc != 0 and b"null pointer" in L.pk_last_error()
    # the Python shim turns the status into RuntimeError with that message
    with pytest.raises(RuntimeError, match="null operand"):
        pk_native._check(L.pk_gemm_tn(0, 0, 8, 8, None, 8, 0, 0, None, 8, 0, 0, None, 8, None, 0, None, 1.0, None, 0, 1, None,
                                      None), "pk_gemm_tn")


def _ref_data_io():
    """The reference's own data_io.py (baseline/_ref, git-ignored copy made by build()) or None."""
    p = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.exists(os.path.join(p, "data_io.py")):
        return None
    import importlib
    sys.path.insert(0, p)
    try:
        for m in ("data_io", "utils"):
            sys.modules.pop(m, None)
        return importlib.import_module("data_io")
    finally:
        sys.path.remove(p)
        sys.modules.pop("utils", None)


def test_text_archives_float_vectors_and_rxspecs_match_the_reference_readers(tmp_path):
    """SURVEY 8f-3 leftovers: ascii matrices (data_io.py:1133-1147), float vectors binary / ascii (:879-990), gzipped files,
    `cmd |` input pipes and :offset suffixes (open_or_fd, :685-718) — against the reference's own readers when available."""
    import gzip
    import pk_train
    rng = np.random.default_rng(3)
    mats = {"utt_a": rng.standard_normal((5, 7)).astype(np.float32), "utt_b": rng.standard_normal((1, 3)).astype(np.float32)}
    txt = tmp_path / "feats.txt.ark"
    with open(txt, "wb") as f:
        for k, m in mats.items():
            f.write((k + "  [\n").encode())
            for i, r in enumerate(m):
                f.write(("  " + " ".join(repr(float(v)) for v in r) + (" ]\n" if i == len(m) - 1 else "\n")).encode())
    got = dict(pk_train.read_mat_ark(open(txt, "rb")))
    assert list(got) == list(mats)
    for k in mats:
        assert np.array_equal(got[k], mats[k])
    vecs = {"v1": rng.standard_normal(6).astype(np.float32), "v2": np.array([], dtype=np.float32), "v3": rng.standard_normal(4)}
    vark = tmp_path / "vec.ark"
    with open(vark, "wb") as f:
        for k, v in vecs.items():
            f.write((k + " ").encode() + b"\0B" + (b"FV " if v.dtype == np.float32 else b"DV ") + b"\x04" +
                    np.int32(v.size).tobytes() + v.tobytes())
        f.write(b"v4 [ 1.25 -2 3e-3 ]\n")
    gv = dict(pk_train.read_vec_flt_ark(open(vark, "rb")))
    for k, v in vecs.items():
        assert np.array_equal(gv[k], v), k
    assert np.allclose(gv["v4"], [1.25, -2.0, 3e-3])
    # binary matrix archive through a gzip file, an input pipe and an scp with byte offsets
    bark = tmp_path / "feats.ark"
    offs = {}
    with open(bark, "wb") as f:
        for k, m in mats.items():
            f.write((k + " ").encode())
            offs[k] = f.tell()
            f.write(b"\0BFM " + b"\x04" + np.int32(m.shape[0]).tobytes() + b"\x04" + np.int32(m.shape[1]).tobytes() + m.tobytes())
    gz = tmp_path / "feats.ark.gz"
    with gzip.open(gz, "wb") as f:
        f.write(open(bark, "rb").read())
    for spec in (f"ark:{gz}", f"ark:cat {bark} |"):
        fd, close = pk_train.open_rx(spec)
        got = dict(pk_train.read_mat_ark(fd))
        if close:
            fd.close()
        assert all(np.array_equal(got[k], mats[k]) for k in mats), spec
    scp = tmp_path / "feats.scp"
    with open(scp, "w") as f:
        for k in mats:
            f.write(f"{k} {bark}:{offs[k]}\n")
    got = dict(pk_train.read_mat_scp(str(scp)))
    assert all(np.array_equal(got[k], mats[k]) for k in mats)
    ref = _ref_data_io()
    if ref is not None:  # the reference's own readers give the same arrays
        out = str(tmp_path)
        rm = {k: np.array(v) for k, v in ref.read_mat_ark(str(txt), out)}
        assert all(np.array_equal(rm[k], mats[k]) for k in mats)
        rv = {k: np.array(v) for k, v in ref.read_vec_flt_ark(str(vark), out)}
        for k in gv:
            assert np.allclose(rv[k], gv[k]), k
        rs = {k: np.array(v) for k, v in ref.read_mat_scp(str(scp), out)}
        assert all(np.array_equal(rs[k], mats[k]) for k in mats)


def test_posterior_and_cntime_readers(tmp_path):
    """SURVEY 8f-3 leftovers: Kaldi `Posterior` archives / scripts and confusion-network bin times
    (data_io.py:1256-1416) against hand-built archives and — when the reference is present — its own readers."""
    import pk_train
    rng = np.random.RandomState(7)
    posts, times = {}, {}
    ark = tmp_path / "post.ark"
    offs = {}
    with open(ark, "wb") as f:
        for k, nfr in (("utt_a", 5), ("utt_b", 1), ("utt_c", 9)):
            frames = []
            f.write((k + " ").encode())
            offs[k] = f.tell()
            f.write(b"\0B\x04" + np.int32(nfr).tobytes())
            for _ in range(nfr):
                n = int(rng.randint(1, 4))
                recs = [(int(rng.randint(0, 1936)), float(np.float32(rng.rand()))) for _ in range(n)]
                f.write(b"\x04" + np.int32(n).tobytes())
                for idx, p in recs:
                    f.write(b"\x04" + np.int32(idx).tobytes() + b"\x04" + np.float32(p).tobytes())
                frames.append(recs)
            posts[k] = frames
    got = dict(pk_train.read_post_ark(open(ark, "rb")))
    assert got == posts
    assert dict(pk_train.read_post_rxspec(f"ark:{ark}")) == posts
    scp = tmp_path / "post.scp"
    with open(scp, "w") as f:
        for k in posts:
            f.write(f"{k} {ark}:{offs[k]}\n")
    assert dict(pk_train.read_post_rxspec(f"scp:{scp}")) == posts
    cark = tmp_path / "cn.ark"
    with open(cark, "wb") as f:
        for k, n in (("utt_a", 4), ("utt_b", 2)):
            t = [(float(np.float32(i * 0.1)), float(np.float32(i * 0.1 + 0.07))) for i in range(n)]
            f.write((k + " ").encode() + b"\0B\x04" + np.int32(n).tobytes())
            for b, e in t:
                f.write(b"\x04" + np.float32(b).tobytes() + b"\x04" + np.float32(e).tobytes())
            times[k] = t
    assert dict(pk_train.read_cntime_ark(open(cark, "rb"))) == times
    with pytest.raises(ValueError):
        list(pk_train.read_post_rxspec("file.ark"))
    ref = _ref_data_io()
    if ref is not None:
        # the reference's archive generators call read_post(fd) / read_cntime(fd) without their second argument
        # (TypeError, data_io.py:1308 / :1377), so its single-entry readers are driven over the archive directly
        out = str(tmp_path)
        for path, one, want in ((ark, ref.read_post, posts), (cark, ref.read_cntime, times)):
            with open(path, "rb") as fd:
                seen = {}
                key = ref.read_key(fd)
                while key:
                    seen[key] = one(fd, out)
                    key = ref.read_key(fd)
            assert seen == want


def test_fusion_rnn_constructor_surface_matches_the_reference():
    """fusionRNN_jit / liGRU_layer / FusionLinearConv (reference :719-995, :2057-2099): same parameter / buffer names,
    shapes and registration order as the reference classes (checkpoints load either way).  The reference hard-codes
    device="cuda" in its constructors; on a CPU-only host `.to("cuda")` and `torch.tensor(..., device="cuda")` are
    patched to no-ops for the duration of ITS construction only.  Values are not compared: the reference moves `u` to
    the GPU before `orthogonal_`, i.e. it draws from the CUDA generator."""
    p = os.path.join(ROOT, "baseline", "_ref", "neural_networks.py")
    if not os.path.exists(p):
        pytest.skip("baseline/_ref missing (python -c 'import __graft_entry__ as g; g.build()')")
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_nn_fusion_cpu", p)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    opts = {"fusionRNN_lay": "16,16,16", "fusionRNN_drop": "0.1,0.1,0.1", "batches": "3", "fusionRNN_do_fusion": "True",
            "fusionRNN_fusion_act": "prelu", "fusionRNN_fusion_reduce": "sum", "fusionRNN_fusion_layer_size": "32",
            "fusionRNN_number_of_mic": "2", "fusionRNN_bidir": "True", "fusionRNN_act": "prelu,prelu,prelu",
            "use_cuda": "True", "to_do": "train"}
    orig_to, orig_tensor = torch.nn.Module.to, torch.tensor

    def to_nocuda(self, *a, **k):
        a = tuple(x for x in a if x != "cuda")
        k = {kk: v for kk, v in k.items() if v != "cuda"}
        return orig_to(self, *a, **k) if (a or k) else self

    def tensor_nocuda(*a, **k):
        if k.get("device") == "cuda":
            k.pop("device")
        return orig_tensor(*a, **k)

    torch.nn.Module.to, torch.tensor = to_nocuda, tensor_nocuda
    try:
        r = ref.fusionRNN_jit(dict(opts), 10)
    finally:
        torch.nn.Module.to, torch.tensor = orig_to, orig_tensor
    m = pknn.fusionRNN_jit(dict(opts), 10)
    assert r.out_dim == m.out_dim == 32
    rs, ms = r.state_dict(), m.state_dict()
    assert list(rs.keys()) == list(ms.keys())
    assert all(rs[k].shape == ms[k].shape and rs[k].dtype == ms[k].dtype for k in rs)
    assert [n for n, _ in r.named_parameters()] == [n for n, _ in m.named_parameters()]
    m.load_state_dict(rs)           # a reference checkpoint loads
    assert m.model[0].do_fusion and not m.model[1].do_fusion and m.model[0].hidden_size == 16
    assert m.model[1].input_size == 32 and m.model[0].wz.in_features == 5 and m.model[0].wz.number_of_mic == 2
    with pytest.raises(RuntimeError):   # no CPU path
        m(torch.zeros(4, 3, 10))


def test_qlstm_constructor_and_weight_assembly_match_the_reference():
    """quaternion_neural_networks.QLSTM / QuaternionLinear(Autograd): same numpy / scipy draws -> identical state_dict; the
    assembled dense matrix equals the reference's Hamilton-product kernel (quaternion_neural_networks.py:375-395)."""
    p = os.path.join(ROOT, "baseline", "_ref", "quaternion_neural_networks.py")
    if not os.path.exists(p):
        pytest.skip("baseline/_ref missing (python -c 'import __graft_entry__ as g; g.build()')")
    import importlib.util
    import quaternion_neural_networks as pkq
    spec = importlib.util.spec_from_file_location("ref_qnn_cpu", p)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    opts = {"lstm_lay": "16,16", "lstm_drop": "0.2,0.2", "lstm_bidir": "True", "lstm_act": "tanh,tanh",
            "quaternion_init": "quaternion", "use_cuda": "False", "to_do": "train"}
    for autograd in ("True", "False"):
        np.random.seed(11)
        torch.manual_seed(11)
        r = ref.QLSTM(dict(opts, autograd=autograd), 12)
        np.random.seed(11)
        torch.manual_seed(11)
        m = pkq.QLSTM(dict(opts, autograd=autograd, use_cuda="True"), 12)
        rs, ms = r.state_dict(), m.state_dict()
        assert list(rs.keys()) == list(ms.keys()) and r.out_dim == m.out_dim == 32
        assert all(torch.equal(rs[k], ms[k]) for k in rs), autograd
        assert [n for n, _ in r.named_parameters()] == [n for n, _ in m.named_parameters()]
        lay_r, lay_m = r.wix[1], m.wix[1]
        x = torch.randn(5, 32)
        want = ref.quaternion_linear(x, lay_r.r_weight, lay_r.i_weight, lay_r.j_weight, lay_r.k_weight, lay_r.bias)
        got = torch.nn.functional.linear(x, lay_m.dense_weight(), lay_m.dense_bias(x.device))
        assert torch.allclose(want, got, rtol=1e-6, atol=1e-6)
    for init in ("unitary", "random"):                       # the other two initialisers, same draws
        np.random.seed(5)
        a = ref.QuaternionLinearAutograd(8, 12, weight_init=init)
        np.random.seed(5)
        b = pkq.QuaternionLinearAutograd(8, 12, weight_init=init)
        assert all(torch.equal(x, y) for x, y in zip(a.state_dict().values(), b.state_dict().values())), init
    with pytest.raises(RuntimeError):                        # no CPU path
        m(torch.zeros(3, 2, 12))


def test_step_entry_point_host_logic(monkeypatch):
    """Host-side mirrors behind pk_rnn_step_* (no kernel runs): which kernel family takes a (cell, H), how many launches a
    call makes, and that the workspace covers the cluster kernels' per-CTA weight images (csrc/pk_cell_cluster*.cu)."""
    pk = pk_native
    monkeypatch.delenv("PK_LSTM_CLUSTER", raising=False)
    monkeypatch.delenv("PK_GRU_CLUSTER", raising=False)
    # defaults: LSTM / GRU / minimalGRU of the shipped recipes run cluster-persistent, one pack + one kernel per call
    for cell in (pk.CELL_LSTM, pk.CELL_GRU, pk.CELL_MGRU):
        assert pk.rnn_step_is_cluster(cell, 550)
        assert pk.rnn_step_launches(cell, 500, 32, 550, 2, False) == pk.rnn_step_launches(cell, 500, 32, 550, 2, True) == 2
    # what they cannot hold falls to the step-wise family
    assert not pk.rnn_step_is_cluster(pk.CELL_LSTM, 600) and not pk.rnn_step_is_cluster(pk.CELL_LIGRU, 2048)
    assert pk.rnn_step_is_cluster(pk.CELL_GRU, 640) and not pk.rnn_step_is_cluster(pk.CELL_GRU, 700)
    assert pk.rnn_step_launches(pk.CELL_LIGRU, 50, 8, 2048, 2, False) == 51        # pack + one launch per step
    # per-CTA images: CL * gates * UPC * (16 KT + 8) halves must fit the workspace (H = 550: 14 x 4 x 40 x 568 x 2 bytes)
    assert pk.rnn_step_workspace_bytes(pk.CELL_LSTM, 500, 32, 550, 2, False) >= 14 * 4 * 40 * 568 * 2
    assert pk.rnn_step_workspace_bytes(pk.CELL_GRU, 500, 32, 550, 2, True) >= 14 * 3 * 40 * 568 * 2
    # the A/B switches
    monkeypatch.setenv("PK_LSTM_CLUSTER", "0")
    monkeypatch.setenv("PK_GRU_CLUSTER", "0")
    assert not pk.rnn_step_is_cluster(pk.CELL_LSTM, 550) and not pk.rnn_step_is_cluster(pk.CELL_MGRU, 550)
    assert pk.rnn_step_launches(pk.CELL_LSTM, 500, 32, 550, 2, True) == 2           # cooperative step-wise kernel + pack
    assert pk.rnn_step_launches(pk.CELL_GRU, 500, 32, 550, 2, True) == 3            # two packs + cooperative kernel
