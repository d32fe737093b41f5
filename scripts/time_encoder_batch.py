import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_model
from augmentedautoencoder_b200 import _lib
enc, _ = make_model(_lib.PREC_TC_SPLIT, 256, 42, with_codebook=False)
for B in (8, 16, 32, 64, 128, 256):
    x = torch.randint(0, 256, (B, 128, 128, 3), dtype=torch.uint8, device="cuda")
    for _ in range(5): enc.encode_device(x)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): enc.encode_device(x)
    b.record(); torch.cuda.synchronize()
    print("B=%3d encoder %.3f ms per call" % (B, a.elapsed_time(b) / 20))
