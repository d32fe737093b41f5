"""Runs a few hot-path steps (config 2: 256 crops, encoder + fused match) for ncu captures.  Not a benchmark."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from augmentedautoencoder_b200 import _lib  # noqa: E402
from augmentedautoencoder_b200.ae.codebook import Codebook  # noqa: E402
from augmentedautoencoder_b200.ae.encoder import Encoder  # noqa: E402
from augmentedautoencoder_b200.ae.session import placeholder  # noqa: E402

prec = _lib.PREC_TC_SPLIT if (len(sys.argv) < 2 or sys.argv[1] == "tc") else _lib.PREC_FP32_SIMT
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
N = 92232
enc = Encoder(placeholder(np.float32, [None, 128, 128, 3]), 128, [128, 256, 512, 512], 5, [2, 2, 2, 2], False, precision=prec, max_batch=256)


class DS:
    embedding_size = N
    _kw = {"num_cyclo": "36"}
    viewsphere_for_embedding = np.zeros((N, 3, 3))


cb = Codebook(enc, DS(), True, max_batch=256, precision=prec)
E = np.random.RandomState(7).standard_normal((N, 128))
cb.embedding_normalized.assign((E / np.linalg.norm(E, axis=1, keepdims=True)).astype(np.float32))
x = torch.randint(0, 256, (256, 128, 128, 3), dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(steps):
    flush.zero_()
    cb.nearest_idx_device(x)
torch.cuda.synchronize()
print("done")
