"""Record the reference constructors' registered state (keys, shapes, order, value checksums under a
fixed seed, generator state after construction) for every class of the module zoo, so the
drop-in's structural parity can be tested without /root/reference (e.g. on the GPU box).

    python tests/golden/make_structure_golden.py
"""
import hashlib
import json
import os
import sys

import torch

sys.path.insert(0, os.environ.get("PK_REFERENCE", "/root/reference"))
import neural_networks as ref_nn  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
from structure_cases import CASES  # noqa: E402


def digest(t):
    return hashlib.sha256(t.detach().contiguous().numpy().tobytes()).hexdigest()[:16]


out = {}
for name, (cls, opts, inp_dim) in CASES.items():
    torch.manual_seed(1234)
    m = getattr(ref_nn, cls)(dict(opts), inp_dim)
    sd = m.state_dict()
    out[name] = dict(
        cls=cls, out_dim=int(m.out_dim),
        keys=[[k, list(v.shape), str(v.dtype), digest(v)] for k, v in sd.items()],
        params=[[k, list(p.shape)] for k, p in m.named_parameters()],
        next_rand=float(torch.rand(1).item()),  # generator position after construction
    )
json.dump(out, open(os.path.join(HERE, "structure.json"), "w"), indent=1)
print("wrote", len(out), "cases")
