"""Device-time the fused codebook match alone (diagnostic, not the benchmark): batch sweep, top-k / upright variants,
optional in-kernel clock trace.

    python scripts/match_bench.py [--rows 92232] [--batches 1,32,128,129,256] [--prec 1] [--k 1] [--upright] [--no-flush]
    AAE_MATCH_TRACE=1 python scripts/match_bench.py --batches 1,256 --iters 3      # per-phase clock64 trace of CTA 0 on stderr

flush (default): a 256 MiB memset and then a 256 MiB read precede every call, so the codebook comes from HBM, the L2 holds
clean lines (a memset alone leaves 126 MB of dirty lines whose write-back shares HBM with the timed kernel: --dirty-flush)
and the queued launch hides the CPU launch latency; --no-flush times back-to-back calls with an idle GPU in between (adds ~12 us of launch latency).
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from augmentedautoencoder_b200.ae.codebook import Codebook  # noqa: E402
from augmentedautoencoder_b200.ae.encoder import Encoder  # noqa: E402
from augmentedautoencoder_b200.ae.session import placeholder  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=92232)
ap.add_argument("--batches", default="1,32,128,129,256")
ap.add_argument("--prec", type=int, default=1)
ap.add_argument("--k", type=int, default=1)
ap.add_argument("--upright", action="store_true")
ap.add_argument("--no-flush", action="store_true")
ap.add_argument("--dirty-flush", action="store_true", help="memset only: leaves 126 MB of dirty lines in L2 whose write-back competes with the timed kernel")
ap.add_argument("--iters", type=int, default=40)
ap.add_argument("--hbm-gbs", type=float, default=6576.1)
ap.add_argument("--floor", action="store_true", help="also time a do-nothing kernel with the match kernel's launch shape (193 KB smem, 512 TMEM columns)")
args = ap.parse_args()

N = args.rows
enc = Encoder(placeholder(np.float32, [None, 128, 128, 3]), 128, [128, 256, 512, 512], 5, [2, 2, 2, 2], False, precision=0, max_batch=256)


class DS:
    embedding_size = N
    _kw = {"num_cyclo": "36"}
    viewsphere_for_embedding = np.zeros((N, 3, 3))


cb = Codebook(enc, DS(), True, max_batch=256, precision=args.prec)
E = np.random.RandomState(7).standard_normal((N, 128))
cb.embedding_normalized.assign((E / np.linalg.norm(E, axis=1, keepdims=True)).astype(np.float32))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
clean = torch.zeros(64 << 20, dtype=torch.int32, device="cuda")          # 256 MiB read after the memset: L2 ends up full of CLEAN lines
if args.floor:
    import ctypes as C
    from augmentedautoencoder_b200 import _lib
    for with_tmem in (0, 1):
        ts = []
        for it in range(args.iters):
            if not args.no_flush:
                flush.zero_()
                if not args.dirty_flush:
                    clean.sum()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _lib.check(_lib.lib().aae_launch_floor_probe(0, with_tmem, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        ts = sorted(ts[min(5, len(ts) - 1):])
        print("launch floor (148 CTAs x 256 threads, 193 KB dynamic smem, tmem alloc=%d, no work): median %.1f us  min %.1f us" % (with_tmem, ts[len(ts) // 2], ts[0]))
for B in [int(b) for b in args.batches.split(",")]:
    z = torch.randn(B, 128, device="cuda")
    ts = []
    for it in range(args.iters):
        if not args.no_flush:
            flush.zero_()
            if not args.dirty_flush:
                clean.sum()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        cb.match_device(z, k=args.k, upright=args.upright)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts = sorted(ts[min(5, len(ts) - 1):])
    med = ts[len(ts) // 2]
    rows = N // 36 if args.upright else N
    byts = rows * 512 + B * 512 + B * args.k * 8
    print("prec=%d k=%d upright=%d B=%3d flush=%d  median %.1f us  min %.1f us   %.2f TB/s = %.2f of %.0f GB/s"
          % (args.prec, args.k, args.upright, B, not args.no_flush, med, ts[0], byts / med / 1e6, byts / med / 1e6 / (args.hbm_gbs / 1e3), args.hbm_gbs))
