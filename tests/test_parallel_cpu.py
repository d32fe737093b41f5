"""CPU tests of the multi-GPU host logic (world_size 2, gloo): shard bounds, top-k all-gather + merge order, latent
all-gather for split encoders, per-object routing.  Device work is replaced by host stand-ins (numpy via the oracle)
through the documented hook methods; the collectives and the partitioning are the real code."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from augmentedautoencoder_b200 import parallel as P  # noqa: E402
from oracle import aae_oracle as O  # noqa: E402


def test_shard_bounds_cover_and_align():
    for n, w, al in [(92232, 8, 36), (368928, 8, 144), (100, 3, 1), (5, 8, 1), (36 * 7, 4, 36)]:
        spans = [P.shard_bounds(n, w, r, al) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        for (a, b), (c, d) in zip(spans, spans[1:]):
            assert b == c and a <= b
        assert all(lo % al == 0 for lo, hi in spans if hi > lo)
    assert P.shard_bounds(368928, 8, 3, 144) == (3 * 46116 + 0, 4 * 46116) or P.shard_bounds(368928, 8, 3, 144)[0] % 144 == 0
    assert [P.split_batch(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert P.owner_of_class([5, 1, 9], 2) == {1: 0, 5: 1, 9: 0}


class HostShard(P.ShardedCodebook):
    """Host stand-in for the CUDA hooks: scores via numpy, merge via a lexicographic sort."""

    def _setup(self):
        self.device = torch.device("cpu")

    def _local_match(self, z, k, upright, s_out, i_out):
        cos = O.cos_similarity(z.numpy(), self._local) if len(self._local) else np.zeros((z.shape[0], 0), np.float32)
        B = z.shape[0]
        s = np.full((B, k), -np.inf, np.float32)
        i = np.full((B, k), -1, np.int32)
        gidx = np.arange(self.lo, self.hi)
        for b in range(B):
            c = cos[b].copy()
            if upright:
                c[gidx % self.num_cyclo != 0] = -np.inf
            order = np.lexsort((gidx, -c))[:k]
            order = order[np.isfinite(c[order])]
            s[b, :len(order)], i[b, :len(order)] = c[order], gidx[order]
        s_out.copy_(torch.from_numpy(s))
        i_out.copy_(torch.from_numpy(i))

    def _merge(self, packed):
        all_s, all_i = packed[:, 0].contiguous().view(torch.float32), packed[:, 1]
        W, B, k = all_s.shape
        so, io = torch.empty((B, k)), torch.empty((B, k), dtype=torch.int32)
        for b in range(B):
            pairs = sorted((-float(all_s[w, b, j]), int(all_i[w, b, j])) for w in range(W) for j in range(k) if int(all_i[w, b, j]) >= 0)[:k]
            for j, (ns, ii) in enumerate(pairs):
                so[b, j], io[b, j] = -ns, ii
        return so, io


class HostRouter(P.ObjectRouter):
    def _run_class(self, cls, crops):
        # deterministic fake "model": score = class id + mean pixel, idx = class id * 1000 + first pixel
        s = crops.float().mean(dim=(1, 2, 3)) + cls
        i = (cls * 1000 + crops[:, 0, 0, 0].int()).to(torch.int32)
        return s[:, None], i[:, None]


def _worker(rank, world, port, fn, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _spawn(fn, world=2):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, fn, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


def _sharded_case(rank, world):
    E = O.make_codebook(3, n=36 * 20, j=16)
    E[36 * 15] = E[36 * 2 + 5]            # the same row in BOTH shards: the lower global index must win
    z = (E[[36 * 15, 7, 36 * 19 + 35, 36 * 11]] * 2.5).astype(np.float32)
    sc = HostShard(E, num_cyclo=36, max_batch=8)
    s, i = sc.match(torch.from_numpy(z), k=3)
    zl = z[slice(*P.split_batch(4, world, rank))]
    s2, i2 = sc.match_split_queries(torch.from_numpy(zl), 4, k=3)
    su, iu = sc.match(torch.from_numpy(z), k=1, upright=True)
    return (sc.lo, sc.hi), s.numpy(), i.numpy(), s2.numpy(), i2.numpy(), iu.numpy()


def test_sharded_match_world2_equals_unsharded():
    out = _spawn(_sharded_case)
    assert out[0][0] == (0, 360) and out[1][0] == (360, 720)
    E = O.make_codebook(3, n=36 * 20, j=16)
    E[36 * 15] = E[36 * 2 + 5]
    z = (E[[36 * 15, 7, 36 * 19 + 35, 36 * 11]] * 2.5).astype(np.float32)
    cos = O.cos_similarity(z, E)
    for r in range(2):
        _, s, i, s2, i2, iu = out[r]
        assert np.array_equal(i[:, 0], np.argmax(cos, axis=1))          # bit-identical to the unsharded argmax
        assert i[0, 0] == 36 * 2 + 5 and i[0, 1] == 36 * 15            # duplicate across shards -> lowest global index first
        assert i[2, 0] == 36 * 19                                       # duplicate cyclo end points inside one shard
        for b in range(4):
            want = np.lexsort((np.arange(720), -cos[b]))[:3]
            assert np.array_equal(i[b], want) and np.allclose(s[b], cos[b, want])
        assert np.array_equal(i2, i) and np.array_equal(s2, s)          # split encoders + latent all-gather: same answer
        assert np.array_equal(iu[:, 0], O.select_indices(cos, upright=True, num_cyclo=36))
    assert np.array_equal(out[0][2], out[1][2])


def _router_case(rank, world):
    classes = [3, 8, 11]
    owner = P.owner_of_class(classes, world)
    mine = {c: None for c, r in owner.items() if r == rank}
    router = HostRouter(mine, classes)
    g = torch.Generator().manual_seed(0)
    crops = torch.randint(0, 256, (10, 4, 4, 3), dtype=torch.uint8, generator=g)
    cls = np.array([3, 8, 11, 3, 99, 8, 8, 11, 3, 3])               # 99: unknown class
    s, i = router.route(crops, cls)
    sh, ih = router.route_host(crops, cls, torch.device("cpu"))       # host routing: only the rank's own crops are touched
    assert torch.equal(s, sh) and torch.equal(i, ih)
    sn, in_ = router.route_host(crops.numpy(), cls, torch.device("cpu"))    # a numpy batch works too
    assert torch.equal(s, sn) and torch.equal(i, in_)
    n_own = sum(len(sel) for _, sel in router.plan(cls))
    assert n_own == sum(1 for c in cls if owner.get(int(c)) == rank) and n_own < len(cls)
    return s.numpy(), i.numpy()


def test_object_routing_world2():
    out = _spawn(_router_case)
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    s, i = out[0]
    g = torch.Generator().manual_seed(0)
    crops = torch.randint(0, 256, (10, 4, 4, 3), dtype=torch.uint8, generator=g)
    cls = [3, 8, 11, 3, 99, 8, 8, 11, 3, 3]
    for b, c in enumerate(cls):
        if c == 99:
            assert i[b] == -1 and s[b] == -np.inf
        else:
            assert i[b] == c * 1000 + int(crops[b, 0, 0, 0]) and abs(s[b] - (crops[b].float().mean().item() + c)) < 1e-4
