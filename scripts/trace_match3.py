import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from augmentedautoencoder_b200.ae.codebook import Codebook
from augmentedautoencoder_b200.ae.encoder import Encoder
from augmentedautoencoder_b200.ae.session import placeholder
N = 92232
enc = Encoder(placeholder(np.float32, [None, 128, 128, 3]), 128, [128, 256, 512, 512], 5, [2, 2, 2, 2], False, precision=0, max_batch=256)
class DS:
    embedding_size = N
    _kw = {"num_cyclo": "36"}
    viewsphere_for_embedding = np.zeros((N, 3, 3))
cbs = []
for seed in (7, 8):
    cb = Codebook(enc, DS(), True, max_batch=256, precision=1)
    E = np.random.RandomState(seed).standard_normal((N, 128))
    cb.embedding_normalized.assign((E / np.linalg.norm(E, axis=1, keepdims=True)).astype(np.float32))
    cbs.append(cb)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
z = torch.randn(256, 128, device="cuda")
for cb in cbs:
    cb.match_device(z)
torch.cuda.synchronize()
for it in range(3):
    flush.zero_()
    cbs[0].match_device(z)   # cold code, cold data
    cbs[1].match_device(z)   # warm code, cold data
    cbs[1].match_device(z)   # warm code, warm data
torch.cuda.synchronize()
