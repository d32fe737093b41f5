import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import aae_oracle as O
from augmentedautoencoder_b200.ae.ae import AE
from augmentedautoencoder_b200.ae.ae_factory import TrainOp
from augmentedautoencoder_b200.ae.decoder import Decoder
from augmentedautoencoder_b200.ae.encoder import Encoder
from augmentedautoencoder_b200.ae.session import placeholder
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
x = placeholder(np.float32, [None, 128, 128, 3]); y = placeholder(np.float32, [None, 128, 128, 3])
enc = Encoder(x, 128, list(O.NUM_FILTER), 5, list(O.STRIDES), False, is_training=True, max_batch=B)
dec = Decoder(y, enc.z, list(reversed(O.NUM_FILTER)), 5, list(reversed(O.STRIDES)), "L2", 4, False, False, is_training=True, max_batch=B)
ep, dp = O.make_encoder_params(42, bias_scale=0.02), O.make_decoder_params(43, bias_scale=0.02)
enc.load_weights(ep); dec.load_weights(dp)
top = TrainOp(AE(enc, dec, 0, 0), 2e-4)
xb = np.random.RandomState(3).rand(B, 128, 128, 3).astype(np.float32)
yb = np.random.RandomState(4).rand(B, 128, 128, 3).astype(np.float32)
loss = top.step_device(torch.from_numpy(xb).cuda(), torch.from_numpy(yb).cuda(), update=False)
loss64, _, g64 = O.ae_forward_loss(xb, yb, ep, dp, dtype=torch.float64, with_grads=True)
_, _, g32 = O.ae_forward_loss(xb, yb, ep, dp, with_grads=True)
print("loss", float(loss), loss64)
grads = top.gradients(torch.device("cuda", 0))
for name, gr in g64.items():
    scale = max(np.abs(gr).max(), 1e-12)
    d = np.abs(grads[name] - gr) / scale
    print("%-18s ours %.2e (n>1e-4: %d of %d)  cpu32 %.2e" % (name, d.max(), int((d > 1e-4).sum()), d.size, np.max(np.abs(g32[name] - gr)) / scale))
