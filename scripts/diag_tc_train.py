"""Tensor-core trainer vs the fp32 CUDA-core trainer (itself pinned to the float64 oracle by tests/test_gpu_a_parity.py):
loss and every gradient after one forward/backward on the same weights and inputs, then step timing at batch 64."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from augmentedautoencoder_b200.ae.ae import AE
from augmentedautoencoder_b200.ae.ae_factory import TrainOp
from augmentedautoencoder_b200.ae.decoder import Decoder
from augmentedautoencoder_b200.ae.encoder import Encoder
from augmentedautoencoder_b200.ae.session import placeholder
from oracle import aae_oracle as O


def build(prec, B, ep, dp):
    x = placeholder(np.float32, [None, 128, 128, 3])
    y = placeholder(np.float32, [None, 128, 128, 3])
    enc = Encoder(x, 128, list(O.NUM_FILTER), 5, list(O.STRIDES), False, is_training=True, max_batch=B, precision=prec)
    dec = Decoder(y, enc.z, list(reversed(O.NUM_FILTER)), 5, list(reversed(O.STRIDES)), "L2", 4, False, False, is_training=True, max_batch=B,
                  precision=prec)
    enc.load_weights(ep)
    dec.load_weights(dp)
    return enc, dec, TrainOp(AE(enc, dec, 0, 0), 2e-4)


def main():
    dev = torch.device("cuda", 0)
    B = int(os.environ.get("DIAG_B", "2"))
    ep, dp = O.make_encoder_params(42, bias_scale=0.02), O.make_decoder_params(43, bias_scale=0.02)
    xb = torch.from_numpy(np.random.RandomState(8).rand(B, 128, 128, 3).astype(np.float32)).cuda()
    yb = torch.from_numpy(np.random.RandomState(4).rand(B, 128, 128, 3).astype(np.float32)).cuda()
    res = {}
    for prec in (0, 1):
        enc, dec, top = build(prec, B, ep, dp)
        loss = float(top.step_device(xb, yb, update=False))
        torch.cuda.synchronize()
        res[prec] = (loss, top.gradients(dev))
        del enc, dec, top
    l0, g0 = res[0]
    l1, g1 = res[1]
    print("loss simt %.9f  tc %.9f  diff %.3e" % (l0, l1, abs(l0 - l1)))
    bad = 0
    for k in g0:
        a, b = g0[k].astype(np.float64), g1[k].astype(np.float64)
        scale = max(np.abs(a).max(), 1e-30)
        rel_l2 = np.linalg.norm(a - b) / max(np.linalg.norm(a), 1e-30)
        mx = np.abs(a - b).max() / scale
        med = np.median(np.abs(a - b)) / scale
        flag = "" if rel_l2 < 2e-3 else "   <-- BAD"
        bad += rel_l2 >= 2e-3
        print("%-22s shape %-22s max|g| %.3e  rel_l2 %.3e  max %.3e  median %.3e%s" % (k, tuple(a.shape), scale, rel_l2, mx, med, flag))
    print("RESULT: %d of %d gradients off" % (bad, len(g0)))
    if os.environ.get("DIAG_TIME", "1") == "1":
        Bt = 64
        xt = torch.rand(Bt, 128, 128, 3, device=dev)
        yt = torch.rand(Bt, 128, 128, 3, device=dev)
        for prec in (0, 1):
            enc, dec, top = build(prec, Bt, ep, dp)
            for _ in range(3):
                top.step_device(xt, yt, update=True)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            n = 10
            for _ in range(n):
                loss = top.step_device(xt, yt, update=True)
            e1.record()
            torch.cuda.synchronize()
            print("precision %d: %.3f ms per training step at batch %d (loss %.6f)" % (prec, e0.elapsed_time(e1) / n, Bt, float(loss)))
            del enc, dec, top


if __name__ == "__main__":
    main()
