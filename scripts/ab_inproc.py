"""In-process A/B of a per-launch environment switch (process-to-process noise on the pool's boxes is +-5 %, more than most kernel
changes): alternates the variable between two values batch by batch and compares the encoder's stage times.
    python scripts/ab_inproc.py AAE_TC_NO_TMA_OUT 0 1 [--batches 60]
Only switches that the library reads on every launch work here (AAE_TC_NO_TMA_OUT, AAE_TC_S5, AAE_TC_EPI8 / AAE_TC_EPI4)."""
import ctypes as C
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from augmentedautoencoder_b200 import _lib  # noqa: E402
from bench import make_model  # noqa: E402

var, v0, v1 = sys.argv[1], sys.argv[2], sys.argv[3]
n = int(sys.argv[sys.argv.index("--batches") + 1]) if "--batches" in sys.argv else 60
enc, _ = make_model(_lib.PREC_TC_SPLIT, 256, 42, with_codebook=False)
x = torch.randint(0, 256, (256, 128, 128, 3), dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
lib = _lib.lib()
h = enc.handle(torch.device("cuda", 0))
for _ in range(10):
    enc.encode_device(x)
buf = (C.c_float * 16)()
lib.aae_encoder_profile(h, 1, None, 0)
res = {v0: [], v1: []}
for i in range(2 * n):
    v = v0 if i % 2 == 0 else v1
    os.environ[var] = v
    flush.zero_()
    enc.encode_device(x)
    torch.cuda.synchronize()
    k = lib.aae_encoder_profile(h, 1, buf, 16)
    res[v].append([buf[j] for j in range(k)])
for v in (v0, v1):
    med = [statistics.median(col) for col in zip(*res[v])]
    print("%s=%s  median stage ms (conv1..conv4, dense): %s  sum %.4f" % (var, v, ["%.4f" % m for m in med], sum(med)))
