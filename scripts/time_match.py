"""Device-time the fused codebook match alone for several batch sizes (diagnostic, not the benchmark)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from augmentedautoencoder_b200 import _lib  # noqa: E402
from augmentedautoencoder_b200.ae.codebook import Codebook  # noqa: E402
from augmentedautoencoder_b200.ae.encoder import Encoder  # noqa: E402
from augmentedautoencoder_b200.ae.session import placeholder  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 92232
for prec in (1, 0):
    enc = Encoder(placeholder(np.float32, [None, 128, 128, 3]), 128, [128, 256, 512, 512], 5, [2, 2, 2, 2], False, precision=0, max_batch=256)

    class DS:
        embedding_size = N
        _kw = {"num_cyclo": "36"}
        viewsphere_for_embedding = np.zeros((N, 3, 3))

    cb = Codebook(enc, DS(), True, max_batch=256, precision=prec)
    E = np.random.RandomState(7).standard_normal((N, 128))
    cb.embedding_normalized.assign((E / np.linalg.norm(E, axis=1, keepdims=True)).astype(np.float32))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for B in (1, 32, 128, 129, 256):
        z = torch.randn(B, 128, device="cuda")
        for do_flush in (False, True):
            ts = []
            for it in range(30):
                if do_flush:
                    flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                cb.match_device(z)
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b) * 1e3)
            ts = sorted(ts[5:])
            print("prec=%d B=%3d flush=%d  median %.1f us  min %.1f us" % (prec, B, do_flush, ts[len(ts) // 2], ts[0]))
