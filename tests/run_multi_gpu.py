"""torchrun entry (one rank per GPU, NCCL): row-sharded codebook + per-object routing against the single-GPU answer.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/run_multi_gpu.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from augmentedautoencoder_b200 import _lib  # noqa: E402
from augmentedautoencoder_b200.parallel import ObjectRouter, ShardedCodebook, owner_of_class, split_batch  # noqa: E402
from oracle import aae_oracle as O  # noqa: E402
from tests.test_gpu_a_parity import _codebook, _enc  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
prec = _lib.PREC_TC_SPLIT
# ---- config 5: fine codebook (368 928 rows) row-sharded, encoder split across ranks, NCCL all-gather of latents + top-k ----
n = 368928
E = O.make_codebook(11, n=n, num_cyclo=144)
p = O.make_encoder_params(42)
enc = _enc(prec, 256, p)
sc = ShardedCodebook(E, num_cyclo=144, max_batch=256, precision=prec)
crops = O.make_crops_u8(1234, 256)
a, e = split_batch(256, world, rank)
z_local = enc.encode_device(torch.from_numpy(crops[a:e]).cuda())
s, i = sc.match_split_queries(z_local, 256, k=1)
full = _codebook(enc, E, num_cyclo=144, max_batch=256, precision=prec)
# the single-GPU answer on the SAME latents: every slice encoded at the batch size its rank uses (below ~64 crops the conv layers
# split K over the idle SM pairs, so a latent's last bits depend on the batch it was encoded in -- like any fp32 conv library)
z_all = torch.cat([enc.encode_device(torch.from_numpy(crops[slice(*split_batch(256, world, r))]).cuda()) for r in range(world)])
s1, i1 = full.match_device(z_all)
assert torch.equal(i, i1) and torch.equal(s, s1), "sharded != unsharded"
z_full = enc.encode_device(torch.from_numpy(crops).cuda())      # one 256-crop batch: same latents up to fp32 rounding of the K split
assert float((z_full - z_all).abs().max()) <= 2e-5 * float(z_full.abs().max()), "slice-encoded latents drifted from the full-batch ones"
# ---- config 4: one object per GPU, mixed batch routed by class ----
classes = list(range(world))
own = owner_of_class(classes, world)
cbs = {}
for c in classes:
    if own[c] == rank:
        enc_c = _enc(prec, 256, O.make_encoder_params(42 + c))
        cbs[c] = _codebook(enc_c, O.make_codebook(7 + c), max_batch=256, precision=prec)
router = ObjectRouter(cbs, classes)
cls = np.random.RandomState(99).randint(0, world, 256)
sr, ir = router.route(torch.from_numpy(crops).cuda(), cls)
mine = np.nonzero(np.array([own[c] for c in cls]) == rank)[0]
for c in cbs:
    sel = np.nonzero(cls == c)[0]
    s_c, i_c = cbs[c].nearest_idx_device(torch.from_numpy(crops[sel]).cuda())
    assert torch.equal(ir[torch.from_numpy(sel).cuda()], i_c[:, 0]) and torch.equal(sr[torch.from_numpy(sel).cuda()], s_c[:, 0])
assert int((ir < 0).sum()) == 0
dist.barrier()
if rank == 0:
    print("multi-gpu ok: world", world, "sharded rows/rank", sc.hi - sc.lo, "routed", len(cls))
dist.destroy_process_group()
