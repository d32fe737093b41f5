"""Throughput of the device training-input pipeline at batch 64 (host draws + packing + H2D of the parameters + two kernels)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from augmentedautoencoder_b200.ae.augment import Augmenter
from tests.test_augment_cpu import TEMPLATE_CODE

B = 64
rng = np.random.RandomState(0)
x = torch.from_numpy(rng.randint(0, 256, (B, 128, 128, 3), dtype=np.uint8)).cuda()
bg = torch.from_numpy(rng.randint(0, 256, (B, 128, 128, 3), dtype=np.uint8)).cuda()
mask = torch.from_numpy(rng.rand(B, 128, 128) > 0.4).cuda()
aug = Augmenter(TEMPLATE_CODE, seed=1)
aug.sigma = 0.8
for _ in range(5):
    aug.augment_device(x, mask, bg)
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 50
for _ in range(n):
    out = aug.augment_device(x, mask, bg)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / n
print("end to end (host draws + pack + upload + kernels): %.3f ms per batch of %d = %.0f images/s" % (dt * 1e3, B, B / dt))
P = aug.sample(B)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
for _ in range(n):
    aug.pack(P)
th = (time.perf_counter() - t0) / n
print("host packing alone: %.3f ms per batch" % (th * 1e3))
