"""Diagnostic: GPU time of n back-to-back fused-match launches after an L2 flush (is there a fixed first-launch cost?)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from augmentedautoencoder_b200.ae.codebook import Codebook  # noqa: E402
from augmentedautoencoder_b200.ae.encoder import Encoder  # noqa: E402
from augmentedautoencoder_b200.ae.session import placeholder  # noqa: E402

N = 92232
enc = Encoder(placeholder(np.float32, [None, 128, 128, 3]), 128, [128, 256, 512, 512], 5, [2, 2, 2, 2], False, precision=0, max_batch=256)


class DS:
    embedding_size = N
    _kw = {"num_cyclo": "36"}
    viewsphere_for_embedding = np.zeros((N, 3, 3))


cb = Codebook(enc, DS(), True, max_batch=256, precision=1)
E = np.random.RandomState(7).standard_normal((N, 128))
cb.embedding_normalized.assign((E / np.linalg.norm(E, axis=1, keepdims=True)).astype(np.float32))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for B in (1, 256):
    z = torch.randn(B, 128, device="cuda")
    for ncall in (1, 2, 3, 5):
        ts = []
        for it in range(20):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(ncall):
                cb.match_device(z)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        ts = sorted(ts[5:])
        print("B=%3d calls=%d  median %.1f us" % (B, ncall, ts[len(ts) // 2]))
