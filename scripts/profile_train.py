"""One warm-up and one profiled training step at batch 64 on the tensor-core trainer (run under ncu for the launch list)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scripts.diag_tc_train import build
from oracle import aae_oracle as O

prec = int(os.environ.get("PREC", "1"))
ep, dp = O.make_encoder_params(42, bias_scale=0.02), O.make_decoder_params(43, bias_scale=0.02)
enc, dec, top = build(prec, 64, ep, dp)
x = torch.rand(64, 128, 128, 3, device="cuda")
y = torch.rand(64, 128, 128, 3, device="cuda")
for _ in range(int(os.environ.get("STEPS", "2"))):
    loss = top.step_device(x, y, update=True)
torch.cuda.synchronize()
print("loss", float(loss))
