"""Training throughput with the input pipeline in the loop: per step draw + pack augmentation parameters on the host, paste /
augment 64 images on the GPU, run the tensor-core training step (BASELINE config 3 + SURVEY 8f N4)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from augmentedautoencoder_b200.ae.augment import Augmenter
from oracle import aae_oracle as O
from scripts.diag_tc_train import build
from tests.test_augment_cpu import TEMPLATE_CODE

B, N = 64, 2048
rng = np.random.RandomState(0)
train_x = torch.from_numpy(rng.randint(0, 256, (N, 128, 128, 3), dtype=np.uint8)).cuda()      # rendered views, resident on the GPU
train_y = train_x.clone()
masks = torch.from_numpy(rng.rand(N, 128, 128) > 0.4).cuda()
bgs = torch.from_numpy(rng.randint(0, 256, (N, 128, 128, 3), dtype=np.uint8)).cuda()
aug = Augmenter(TEMPLATE_CODE, seed=1)
ep, dp = O.make_encoder_params(42, bias_scale=0.02), O.make_decoder_params(43, bias_scale=0.02)
enc, dec, top = build(1, B, ep, dp)


def step():
    idx = torch.from_numpy(np.random.choice(N, B, replace=False)).cuda()
    idb = torch.from_numpy(np.random.choice(N, B, replace=False)).cuda()
    x = aug.augment_device(train_x[idx], masks[idx], bgs[idb])
    y = train_y[idx].to(torch.float32) / 255.0
    return top.step_device(x, y, update=True)


for _ in range(5):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 50
for _ in range(n):
    loss = step()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / n
print("augment + train step: %.3f ms per step of %d images = %.0f images/s (loss %.5f)" % (dt * 1e3, B, B / dt, float(loss)))
