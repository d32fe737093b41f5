#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE's own Python host code.

Run in the build container only (needs /root/reference; the GPU box has no copy):

    python tests/golden/make_golden.py

TensorFlow / progressbar / m3vision are not installable here, so they are replaced by inert stubs
before importing ``auto_pose``.  Everything recorded below is computed by *reference* code:

* view sphere table     Dataset.viewsphere_for_embedding              (auto_pose/ae/dataset.py:39-58)
* index selection       Codebook.nearest_rotation (argmax/upright/top_n) (auto_pose/ae/codebook.py:55-75)
                        with ``session.run`` answered by a fake session that returns a fixed cosine matrix
* pose lift             Codebook.auto_pose6d                          (auto_pose/ae/codebook.py:79-129)
* crop extraction       AePoseEstimator.extract_square_patch / process (auto_pose/m3_interface/ae_pose_estimator.py:106-232)
                        Dataset.extract_square_patch                  (auto_pose/ae/dataset.py:354-373)
* preprocessing         the uint8 -> x/255. -> float32 feed           (auto_pose/ae/codebook.py:58-61)

What can NOT be generated (TensorFlow kernels: conv/dense/l2_normalize/matmul/top_k/Adam) stays
"parity unpinned" -- see oracle/aae_oracle.py header.
"""
import configparser
import os
import sys
import types

import numpy as np

sys.dont_write_bytecode = True  # /root/reference is read-only
REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


# ---------------------------------------------------------------------------------------- stubs
class _Anything(types.ModuleType):
    """Module/object stub: any attribute access or call yields another stub."""

    def __init__(self, name="stub"):
        super().__init__(name)

    def __getattr__(self, item):
        if item.startswith("__") and item.endswith("__"):
            raise AttributeError(item)
        return _Anything(f"{self.__name__}.{item}")

    def __call__(self, *a, **k):
        return _Anything(self.__name__ + "()")

    def __iter__(self):
        return iter(())


def install_stubs():
    tf = _Anything("tensorflow")
    sys.modules["tensorflow"] = tf
    sys.modules["tensorflow.compat"] = _Anything("tensorflow.compat")
    sys.modules["tensorflow.compat.v1"] = tf
    sys.modules["progressbar"] = _Anything("progressbar")
    sys.modules["tf_slim"] = _Anything("tf_slim")
    if not hasattr(np, "float"):
        np.float = float  # auto_pose/ae/dataset.py:35 uses the removed alias
    sys.path.insert(0, REF)
    # m3vision is external to the reference; the reference vendors the same ABCs in m3_interfaces.py
    import importlib.util
    spec = importlib.util.spec_from_file_location("m3_interfaces", f"{REF}/auto_pose/m3_interface/m3_interfaces.py")
    m3i = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m3i)
    m3 = types.ModuleType("m3vision")
    m3.interfaces = types.ModuleType("m3vision.interfaces")
    m3.interfaces.pose_estimator = m3i
    sys.modules["m3vision"] = m3
    sys.modules["m3vision.interfaces"] = m3.interfaces
    sys.modules["m3vision.interfaces.pose_estimator"] = m3i
    return m3i


class FakeSession:
    """Stands in for tf.Session: answers .run(fetch, feed) from a table keyed by the fetch object."""

    def __init__(self):
        self.table = {}
        self.last_feed = None

    def run(self, fetch, feed_dict=None):
        self.last_feed = feed_dict
        v = self.table[id(fetch)]
        return v(feed_dict) if callable(v) else v


def template_dataset_kw():
    cfg = configparser.ConfigParser()
    cfg.read(f"{REF}/auto_pose/ae/cfg/train_template.cfg")
    kw = {}
    for sec in ("Dataset", "Paths", "Augmentation", "Queue", "Embedding"):
        kw.update({k: v for k, v in cfg.items(sec)})
    return cfg, kw


def main():
    m3i = install_stubs()
    from auto_pose.ae.dataset import Dataset
    from auto_pose.ae.codebook import Codebook

    cfg, kw = template_dataset_kw()
    kw["noof_training_imgs"] = "1"
    kw["noof_bg_imgs"] = "1"
    kw["background_images_glob"] = "/nonexistent/*.jpg"
    ds = Dataset("/tmp/unused", **kw)

    # ---- 1. view sphere ------------------------------------------------------------------
    rs = ds.viewsphere_for_embedding
    assert rs.shape == (92232, 3, 3)
    np.savez_compressed(f"{OUT}/viewsphere_2562x36.npz", view_R=rs[::36].copy(), first_rows=rs[:72].copy(),
                        last_rows=rs[-36:].copy(), checksum=np.array([rs.sum(), np.abs(rs).sum()]),
                        probe_idx=np.arange(0, 92232, 4099), probe_R=rs[::4099].copy())
    # a small sphere for cheap full comparisons
    kw2 = dict(kw)
    kw2["min_n_views"], kw2["num_cyclo"] = "162", "12"
    ds_small = Dataset("/tmp/unused", **kw2)
    np.savez_compressed(f"{OUT}/viewsphere_162x12.npz", R=ds_small.viewsphere_for_embedding)

    # ---- 2. index selection + preprocessing through Codebook.nearest_rotation ---------------
    class Enc:  # minimal stand-in for the Encoder object the Codebook wraps
        latent_space_size = 128
        x = object()
        z = object()

    cb = Codebook(Enc(), ds, True)
    sess = FakeSession()
    rng = np.random.RandomState(5)
    n = ds.embedding_size
    cos_b = rng.standard_normal((6, n)).astype(np.float32)
    cos_b[1, 777] = cos_b[1, 40000] = cos_b[1].max() + 1.0          # exact tie -> lowest index
    cos_b[2, 36 * 100 + 35] = cos_b[2, 36 * 100] = cos_b[2].max() + 2.0  # duplicate cyclo end points
    cos_b[3, 5] = cos_b[3].max() + 3.0                                # best is not a multiple of 36
    sess.table[id(cb.cos_similarity)] = cos_b
    x_u8 = rng.randint(0, 256, size=(6, 128, 128, 3), dtype=np.uint8)
    idc_plain = cb.nearest_rotation(sess, x_u8, return_idcs=True)
    fed = sess.last_feed[Enc.x]  # what the reference would hand to the float32 placeholder
    idc_upright = cb.nearest_rotation(sess, x_u8, upright=True, return_idcs=True)
    r_batch = cb.nearest_rotation(sess, x_u8)
    sess.table[id(cb.cos_similarity)] = cos_b[:1]
    r_single = cb.nearest_rotation(sess, x_u8[0])
    idc_top8 = cb.nearest_rotation(sess, x_u8[0], top_n=8, return_idcs=True)
    sess.table[id(cb.nearest_neighbor_idx)] = np.argmax(cos_b, axis=1)
    r_nn_batch = cb.nearest_rotation_batch(sess, x_u8)
    np.savez_compressed(
        f"{OUT}/select.npz", cos_seed=np.array(5), tie_rows=np.array([[1, 777, 40000], [2, 3600, 3635], [3, 5, 5]]),
        idc_plain=idc_plain, idc_upright=idc_upright, r_batch=r_batch, r_single=r_single, idc_top8=idc_top8,
        r_nn_batch=r_nn_batch, fed_dtype=np.array(str(fed.dtype)), x_probe=x_u8[0, :2, :4],
        fed_probe=np.asarray(fed, dtype=np.float32)[0, :2, :4],
        u8_over_255_f32=(np.arange(256, dtype=np.uint8) / 255.).astype(np.float32))

    # ---- 3. pose lift (auto_pose6d) -----------------------------------------------------------
    bbs = np.zeros((n, 4), dtype=np.int32)
    bbs[:, 0] = rng.randint(200, 400, n)
    bbs[:, 1] = rng.randint(100, 300, n)
    bbs[:, 2] = rng.randint(60, 200, n)
    bbs[:, 3] = rng.randint(60, 200, n)
    cb.embed_obj_bbs_values = bbs
    k_test = np.array([[572.4114, 0, 325.2611], [0, 573.57043, 242.04899], [0, 0, 1]])
    cases = []
    for q, (bb, topn, upright) in enumerate([([250.0, 120.0, 88.0, 140.0], 1, False), ([10.0, 300.0, 40.5, 30.25], 1, True),
                                             ([400.0, 50.0, 120.0, 60.0], 4, False)]):
        sess.table[id(cb.cos_similarity)] = cos_b[q:q + 1]
        rs_est, ts_est = cb.auto_pose6d(sess, x_u8[q], bb, k_test, topn, cfg, upright=upright)
        idcs = cb.nearest_rotation(sess, x_u8[q], top_n=topn, upright=upright, return_idcs=True)
        cases.append((np.array(bb), topn, upright, np.array(idcs), rs_est, ts_est))
    rs_d, ts_d = cb.auto_pose6d(sess, x_u8[2], [400.0, 50.0, 120.0, 60.0], k_test, 1, cfg, depth_pred=812.5)
    np.savez_compressed(
        f"{OUT}/pose_lift.npz", bbs_seed_note=np.array("RandomState(5) continued; bbs stored sparsely"),
        k_test=k_test, k_train=np.array(eval(cfg.get("Dataset", "K"))).reshape(3, 3),
        radius=np.array(cfg.getfloat("Dataset", "RADIUS")),
        **{f"c{i}_bb": c[0] for i, c in enumerate(cases)}, **{f"c{i}_topn": np.array(c[1]) for i, c in enumerate(cases)},
        **{f"c{i}_upright": np.array(c[2]) for i, c in enumerate(cases)}, **{f"c{i}_idcs": c[3] for i, c in enumerate(cases)},
        **{f"c{i}_bbs_at_idcs": bbs[np.atleast_1d(c[3])] for i, c in enumerate(cases)},
        **{f"c{i}_Rs": c[4] for i, c in enumerate(cases)}, **{f"c{i}_ts": c[5] for i, c in enumerate(cases)},
        depth_Rs=rs_d, depth_ts=ts_d, depth_idx=np.atleast_1d(np.argmax(cos_b[2])), depth_bb=bbs[np.argmax(cos_b[2])])

    # ---- 4. crops + AePoseEstimator.process ---------------------------------------------------
    from auto_pose.m3_interface.ae_pose_estimator import AePoseEstimator
    est = AePoseEstimator.__new__(AePoseEstimator)  # __init__ needs a TF session + checkpoints
    scene = np.random.RandomState(11).randint(0, 256, size=(480, 640, 3), dtype=np.uint8)  # not stored: re-derived from the seed
    import cv2
    boxes = [[100.3, 80.7, 120.2, 90.9], [5.0, 5.0, 60.0, 200.0], [500.0, 300.0, 139.0, 179.0]]
    crops_bb = np.stack([est.extract_square_patch(scene, b, 1.2, resize=(128, 128), interpolation=cv2.INTER_LINEAR,
                                                  black_borders=True) for b in boxes])
    crops_ds = np.stack([ds.extract_square_patch(scene, b, 1.2, resize=(128, 128), interpolation=cv2.INTER_NEAREST)
                         for b in boxes])
    # full process(): wire the estimator by hand around the fake session
    est._camPose, est._upright, est._topk = False, False, 1
    est.class_2_encoder = {1: "grp/exp"}
    est.all_codebooks = {1: cb}
    est.all_train_args = {1: cfg}
    est.pad_factors = {1: 1.2}
    est.patch_sizes = {1: (128, 128)}
    est.sess = sess
    fed_crops = []
    cos_rows = iter([cos_b[4:5], cos_b[5:6]])

    def answer(feed):
        fed_crops.append(np.asarray(feed[Enc.x], dtype=np.float32))
        return next(cos_rows)

    sess.table[id(cb.cos_similarity)] = answer
    dets = [m3i.BoundingBox(xmin=0.2, ymin=0.25, xmax=0.45, ymax=0.6, classes={1: 0.9, 2: 0.1}),
            m3i.BoundingBox(xmin=0.5, ymin=0.1, xmax=0.9, ymax=0.5, classes={7: 0.9}),          # unknown class: skipped
            m3i.BoundingBox(xmin=0.6, ymin=0.5, xmax=0.95, ymax=0.9, classes={1: 0.8})]
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        poses = est.process(dets, scene, k_test, mm=False)
    fed = np.concatenate(fed_crops)
    fed_u8 = np.rint(fed * 255.0).astype(np.uint8)  # the feed is exactly u8/255. -> store the u8 form
    assert np.array_equal((fed_u8 / 255.).astype(np.float32), fed)
    np.savez_compressed(
        f"{OUT}/crops_process.npz", scene_seed=np.array(11), scene_shape=np.array(scene.shape),
        boxes=np.array(boxes), crops_black_borders_linear=crops_bb, crops_dataset_nearest=crops_ds,
        det_boxes=np.array([[0.2, 0.25, 0.45, 0.6], [0.5, 0.1, 0.9, 0.5], [0.6, 0.5, 0.95, 0.9]]),
        det_classes=np.array([1, 7, 1]), fed_crops_u8=fed_u8, cos_rows_used=np.array([4, 5]),
        pose_names=np.array([p.name for p in poses]), pose_trafos=np.stack([p.trafo for p in poses]),
        bbs_at_best=bbs[[int(np.argmax(cos_b[4])), int(np.argmax(cos_b[5]))]],
        best_idx=np.array([int(np.argmax(cos_b[4])), int(np.argmax(cos_b[5]))]))
    print("golden fixtures written to", OUT)
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f"  {f}: {os.path.getsize(os.path.join(OUT, f))} B")


if __name__ == "__main__":
    main()
