"""Device-time Decoder.x (batch 64) on the fp32 SIMT and tensor-core paths (diagnostic)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from augmentedautoencoder_b200.ae.decoder import Decoder
from augmentedautoencoder_b200.ae.session import placeholder
B = 64
z = torch.randn(B, 128, device="cuda")
for prec in (0, 1):
    dec = Decoder(placeholder(np.float32, [None, 128, 128, 3]), placeholder(np.float32, [None, 128]), [512, 512, 256, 128], 5, [2, 2, 2, 2], "L2", 4,
                  False, False, max_batch=B, precision=prec)
    for _ in range(3):
        dec.decode_device(z)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        dec.decode_device(z)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print("precision %d: %.3f ms per decode of %d latents (%.1f useful TFLOP/s on 17.1 GFLOP/image reference count)" % (prec, ms, B, 17.1002e9 * B / ms / 1e9))
