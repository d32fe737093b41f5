"""Multi-GPU partitioning of the hot path: one process per GPU, torch.distributed (NCCL over NVLink / NVSwitch) for the
plumbing.  The reference has no multi-device code at all ("Parallelism strategies: none"); these are
the three ways the path shards naturally:

* query data-parallel  -- crops are independent: every rank holds a replica of (encoder, codebook) and its slice of the
                          batch; no data-path collective (``split_batch``).
* one object per GPU   -- an object class is an independent (encoder, codebook) pair (auto_pose/m3_interface/
                          ae_pose_estimator.py:48-78): crops are routed to the rank that owns their class (``ObjectRouter``).
* row-sharded codebook -- a single large codebook split by rows: rank r scores all queries against rows
                          [lo_r, hi_r), the per-shard top-k (score, global index) lists are all-gathered (B*k*8 bytes per
                          rank: pure latency) and merged with "highest score, then LOWEST global index", which makes the
                          result bit-identical to the unsharded np.argmax of auto_pose/ae/codebook.py:63-68
                          (``ShardedCodebook``).
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib


# --------------------------------------------------------------------------------------------------------- partitioning
def shard_bounds(n_rows, world_size, rank, align=1):
    """Row range [lo, hi) of `rank` when n_rows are split as evenly as possible into world_size contiguous shards
    (shard starts aligned to `align`, e.g. NUM_CYCLO so that an `upright` search never straddles shards)."""
    per = -(-n_rows // world_size)
    per = -(-per // align) * align
    lo = min(rank * per, n_rows)
    hi = min(lo + per, n_rows)
    return lo, hi


def split_batch(batch, world_size, rank):
    """Contiguous slice [a, e) of a query batch for data-parallel replicas."""
    per = -(-batch // world_size)
    a = min(rank * per, batch)
    return a, min(a + per, batch)


def owner_of_class(class_ids, world_size):
    """class id -> owning rank: round-robin over the sorted class list (deterministic on every rank)."""
    return {c: i % world_size for i, c in enumerate(sorted(class_ids))}


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


# --------------------------------------------------------------------------------------------------------- sharded codebook
class ShardedCodebook:
    """Row shard `rank` of a codebook + the all-gather/merge that reconstitutes the global top-k.

    embedding: the FULL [N, J] float32 table (every rank slices its own rows; a loader that only reads its rows can pass
    ``row_range`` and the slice instead)."""

    def __init__(self, embedding, num_cyclo=36, max_batch=256, precision=None, device=None, group=None,
                 row_range=None, n_rows_total=None):
        self.group = group
        self.rank, self.world = _world(group)
        emb = np.asarray(embedding, dtype=np.float32)
        if row_range is None:
            self.n_total = emb.shape[0]
            self.lo, self.hi = shard_bounds(self.n_total, self.world, self.rank, align=int(num_cyclo))
            emb = emb[self.lo:self.hi]
        else:
            self.lo, self.hi = row_range
            self.n_total = int(n_rows_total)
            assert emb.shape[0] == self.hi - self.lo
        self.latent = emb.shape[1]
        # precision=None: the tensor-core match where the latent size allows it, else the fp32 kernels (as ae.codebook.Codebook)
        self._auto_precision = precision is None
        self.num_cyclo, self.max_batch = int(num_cyclo), int(max_batch)
        self.precision = _lib.PREC_TC_SPLIT if precision is None else int(precision)
        self.device = device
        self._local = np.ascontiguousarray(emb)
        self._handle = None
        self._setup()

    # -- device hooks (overridden by the CPU/gloo tests with host stand-ins) -----------------------------------------
    def _setup(self):
        if self.device is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        if self._local.shape[0] == 0:
            return
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            st = _lib.lib().aae_codebook_create(self.device.index, _lib.ptr(self._local), self._local.shape[0], self.latent, self.num_cyclo, self.lo,
                                                self.max_batch, self.precision, C.byref(h))
            if st == -3 and self._auto_precision and self.precision == _lib.PREC_TC_SPLIT:
                self.precision = _lib.PREC_FP32_SIMT
                st = _lib.lib().aae_codebook_create(self.device.index, _lib.ptr(self._local), self._local.shape[0], self.latent, self.num_cyclo,
                                                    self.lo, self.max_batch, self.precision, C.byref(h))
            _lib.check(st, "sharded codebook create")
        self._handle = h

    def _local_match(self, z, k, upright):
        """z [B, J] on self.device -> (scores [B, k], global idx [B, k]); an empty shard returns (-inf, -1)."""
        B = z.shape[0]
        scores = torch.full((B, k), float("-inf"), dtype=torch.float32, device=z.device)
        idx = torch.full((B, k), -1, dtype=torch.int32, device=z.device)
        if self._handle is None:
            return scores, idx
        kk = min(k, self.hi - self.lo)
        s_loc = torch.empty((B, kk), dtype=torch.float32, device=z.device)
        i_loc = torch.empty((B, kk), dtype=torch.int32, device=z.device)
        stream = C.c_void_p(torch.cuda.current_stream(z.device).cuda_stream)
        for a in range(0, B, self.max_batch):
            e = min(B, a + self.max_batch)
            _lib.check(_lib.lib().aae_codebook_match(self._handle, _lib.ptr(z[a:e]), e - a, kk, int(bool(upright)), _lib.ptr(s_loc[a:e]),
                                                     _lib.ptr(i_loc[a:e]), stream), "sharded match")
        scores[:, :kk], idx[:, :kk] = s_loc, i_loc
        return scores, idx

    def _merge(self, all_scores, all_idx):
        """[W, B, k] gathered lists -> [B, k]: score descending, ties to the lowest global index (aae_topk_merge)."""
        W, B, k = all_scores.shape
        so = torch.empty((B, k), dtype=torch.float32, device=all_scores.device)
        io = torch.empty((B, k), dtype=torch.int32, device=all_scores.device)
        _lib.check(_lib.lib().aae_topk_merge(_lib.ptr(all_scores), _lib.ptr(all_idx), W, B, k, _lib.ptr(so), _lib.ptr(io),
                                             C.c_void_p(torch.cuda.current_stream(all_scores.device).cuda_stream)), "topk merge")
        return so, io

    # -- the exchange step -------------------------------------------------------------------------------------------
    def match(self, z, k=1, upright=False):
        """Every rank passes the same queries z [B, J]; every rank gets the global (scores [B,k], idx [B,k])."""
        s, i = self._local_match(z.contiguous(), k, upright)
        if self.world == 1:
            return s, i
        B = s.shape[0]
        all_s = torch.empty((self.world * B, k), dtype=s.dtype, device=s.device)   # rank-major concatenation = [W, B, k]
        all_i = torch.empty((self.world * B, k), dtype=i.dtype, device=i.device)
        dist.all_gather_into_tensor(all_s, s.contiguous(), group=self.group)
        dist.all_gather_into_tensor(all_i, i.contiguous(), group=self.group)
        return self._merge(all_s.view(self.world, B, k), all_i.view(self.world, B, k))

    def match_split_queries(self, z_local, batch_total, k=1, upright=False):
        """Encoder work split across ranks: each rank encoded only its ``split_batch`` slice; the latents ([B_r, J],
        512 B per query) are all-gathered first, then matched as in ``match``."""
        if self.world == 1:
            return self.match(z_local, k, upright)
        per = -(-batch_total // self.world)
        pad = torch.zeros((per, z_local.shape[1]), dtype=z_local.dtype, device=z_local.device)
        pad[:z_local.shape[0]] = z_local
        all_z = torch.empty((self.world * per, z_local.shape[1]), dtype=z_local.dtype, device=z_local.device)
        dist.all_gather_into_tensor(all_z, pad, group=self.group)
        return self.match(all_z[:batch_total].contiguous(), k, upright)

    def close(self):
        if self._handle is not None:
            _lib.lib().aae_codebook_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# --------------------------------------------------------------------------------------------------------- object routing
class ObjectRouter:
    """One (encoder, codebook) pair per object class, classes spread over the ranks (BASELINE config 4).  Every rank sees
    the same mixed batch of crops and class ids; it runs the crops of the classes it owns, and the per-crop results are
    combined with a MAX all-reduce (every position is written by exactly one rank, all others hold the identity)."""

    def __init__(self, codebooks_by_class, all_class_ids, group=None):
        """codebooks_by_class: {class id: Codebook} for the classes THIS rank owns (see ``owner_of_class``)."""
        self.group = group
        self.rank, self.world = _world(group)
        self.owner = owner_of_class(all_class_ids, self.world)
        self.codebooks = dict(codebooks_by_class)
        missing = [c for c, r in self.owner.items() if r == self.rank and c not in self.codebooks]
        if missing:
            raise ValueError("rank %d owns classes %s but has no codebook for them" % (self.rank, missing))

    def _run_class(self, cls, crops_dev):
        return self.codebooks[cls].nearest_idx_device(crops_dev, k=1)

    def route(self, crops_dev, class_ids):
        """crops_dev [B,H,W,C] on this rank's device, class_ids: length-B sequence.  Returns (scores [B], idx [B]) complete on
        every rank; crops of unknown classes get (-inf, -1) (the reference skips them, ae_pose_estimator.py:147-149)."""
        class_ids = np.asarray(class_ids)
        B = len(class_ids)
        scores = torch.full((B,), float("-inf"), dtype=torch.float32, device=crops_dev.device)
        idx = torch.full((B,), -1, dtype=torch.int32, device=crops_dev.device)
        for cls, owner in self.owner.items():
            if owner != self.rank:
                continue
            sel = np.nonzero(class_ids == cls)[0]
            if len(sel) == 0:
                continue
            sel_t = torch.from_numpy(sel).to(crops_dev.device)
            s, i = self._run_class(cls, crops_dev.index_select(0, sel_t).contiguous())
            scores[sel_t] = s[:, 0]
            idx[sel_t] = i[:, 0]
        if self.world > 1:
            dist.all_reduce(scores, op=dist.ReduceOp.MAX, group=self.group)
            dist.all_reduce(idx, op=dist.ReduceOp.MAX, group=self.group)
        return scores, idx
