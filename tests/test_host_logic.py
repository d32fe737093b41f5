"""CPU tests of the product's host-side logic against the golden vectors produced by the reference's own code
(tests/golden/make_golden.py): view-sphere table, pose lift, crop extraction, checkpoint round trip, session shim."""
import os

import numpy as np
import pytest

from augmentedautoencoder_b200.ae import utils
from augmentedautoencoder_b200.ae.codebook import lift_pose, lift_pose_batch
from augmentedautoencoder_b200.ae.dataset import Dataset


def g(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def test_viewsphere_matches_reference(golden_dir):
    small = Dataset(None, min_n_views=162, num_cyclo=12, radius=700)
    assert np.array_equal(small.viewsphere_for_embedding, g(golden_dir, "viewsphere_162x12.npz")["R"])
    full = Dataset(None, min_n_views=2562, num_cyclo=36, radius=700)
    ref = g(golden_dir, "viewsphere_2562x36.npz")
    R = full.viewsphere_for_embedding
    assert full.embedding_size == 92232
    assert np.array_equal(R[::36], ref["view_R"]) and np.array_equal(R[:72], ref["first_rows"])
    assert np.array_equal(R[-36:], ref["last_rows"]) and np.array_equal(R[ref["probe_idx"]], ref["probe_R"])


def test_lift_pose_batch_is_bit_identical_to_the_per_detection_call(golden_dir):
    """the vectorised pose lift used by AePoseEstimator.process == the reference's per-detection tail
    (codebook.py:82-129), bit for bit, on the reference-generated golden and on 3000 random detections."""
    p = g(golden_dir, "pose_lift.npz")
    rs = Dataset(None, min_n_views=2562, num_cyclo=36, radius=700).viewsphere_for_embedding
    for i in range(3):
        idcs = np.atleast_1d(p[f"c{i}_idcs"])
        table = np.zeros((92232, 4), dtype=np.int32)
        table[idcs] = p[f"c{i}_bbs_at_idcs"]
        r, t = lift_pose_batch(idcs[None, :], rs, table, np.asarray(p[f"c{i}_bb"])[None], p["k_test"], p["k_train"], float(p["radius"]))
        assert np.array_equal(r[0], p[f"c{i}_Rs"]) and np.array_equal(t[0], p[f"c{i}_ts"])
    table = np.zeros((92232, 4), dtype=np.int32)
    table[int(p["depth_idx"][0])] = p["depth_bb"]
    r, t = lift_pose_batch(np.asarray(p["depth_idx"])[None, :], rs, table, np.asarray(p["c2_bb"])[None], p["k_test"], p["k_train"],
                           float(p["radius"]), depth_pred=812.5)
    assert np.array_equal(r[0], p["depth_Rs"]) and np.array_equal(t[0], p["depth_ts"])
    rng = np.random.RandomState(0)
    D, k = 3000, 3
    bbs = rng.randint(30, 500, (92232, 4)).astype(np.int32)
    idcs = rng.randint(0, 92232, (D, k))
    pb = np.stack([rng.uniform(0, 500, D), rng.uniform(0, 400, D), rng.uniform(20, 300, D), rng.uniform(20, 300, D)], 1)
    pb[::3] = np.rint(pb[::3])                       # detectors mostly deliver integer pixel boxes
    Rb, tb = lift_pose_batch(idcs, rs, bbs, pb, p["k_test"], p["k_train"], 700.0)
    assert Rb.shape == (D, k, 3, 3) and tb.shape == (D, k, 3) and Rb.dtype == np.float64 and tb.dtype == np.float64
    for d in range(D):
        Rl, tl = lift_pose(idcs[d], rs, bbs, pb[d], p["k_test"], p["k_train"], 700.0)
        assert np.array_equal(Rl, Rb[d]) and np.array_equal(tl, tb[d]), d


def test_lift_pose_matches_reference(golden_dir):
    p = g(golden_dir, "pose_lift.npz")
    rs = Dataset(None, min_n_views=2562, num_cyclo=36, radius=700).viewsphere_for_embedding
    for i in range(3):
        idcs = np.atleast_1d(p[f"c{i}_idcs"])
        table = np.zeros((92232, 4), dtype=np.int32)
        table[idcs] = p[f"c{i}_bbs_at_idcs"]
        r, t = lift_pose(idcs, rs, table, p[f"c{i}_bb"], p["k_test"], p["k_train"], float(p["radius"]))
        assert np.array_equal(r, p[f"c{i}_Rs"]) and np.array_equal(t, p[f"c{i}_ts"])
    table = np.zeros((92232, 4), dtype=np.int32)
    table[int(p["depth_idx"][0])] = p["depth_bb"]
    r, t = lift_pose(p["depth_idx"], rs, table, p["c2_bb"], p["k_test"], p["k_train"], float(p["radius"]), depth_pred=812.5)
    assert np.array_equal(r, p["depth_Rs"]) and np.array_equal(t, p["depth_ts"])


def test_square_patches_match_reference(golden_dir):
    import cv2
    from augmentedautoencoder_b200.m3_interface.ae_pose_estimator import AePoseEstimator
    c = g(golden_dir, "crops_process.npz")
    scene = np.random.RandomState(int(c["scene_seed"])).randint(0, 256, size=tuple(c["scene_shape"]), dtype=np.uint8)
    est = AePoseEstimator.__new__(AePoseEstimator)
    for b, want in zip(c["boxes"], c["crops_black_borders_linear"]):
        got = est.extract_square_patch(scene, b, 1.2, resize=(128, 128), interpolation=cv2.INTER_LINEAR, black_borders=True)
        assert np.array_equal(got, want)
    ds = Dataset(None)
    for b, want in zip(c["boxes"], c["crops_dataset_nearest"]):
        assert np.array_equal(ds.extract_square_patch(scene, b, 1.2, resize=(128, 128), interpolation=cv2.INTER_NEAREST), want)


def test_workspace_paths_and_batches():
    assert utils.get_log_dir("/w", "exp", "grp") == "/w/experiments/grp/exp"
    assert utils.get_checkpoint_dir("/w/experiments/grp/exp") == "/w/experiments/grp/exp/checkpoints"
    assert utils.get_train_config_exp_file_path("/l", "exp") == "/l/exp.cfg"
    assert list(utils.batch_iteration_indices(10, 4)) == [(0, 4), (4, 8), (8, 10)]
    assert list(utils.batch_iteration_indices(8, 4)) == [(0, 4), (4, 8)]


def test_variable_scope_names_follow_the_reference():
    from augmentedautoencoder_b200 import _lib
    from augmentedautoencoder_b200.ae import session as S
    from augmentedautoencoder_b200.ae.encoder import Encoder
    _lib.lib()
    with S.variable_scope("obj_05"):
        x = S.placeholder(np.float32, [None, 128, 128, 3])
        enc = Encoder(x, 128, [128, 256, 512, 512], 5, [2, 2, 2, 2], False)
    names = enc.variable_names
    assert names[0] == "obj_05/conv2d/kernel" and names[2] == "obj_05/conv2d_1/kernel" and names[-2:] == ["obj_05/dense/kernel", "obj_05/dense/bias"]
    w = enc.get_weights()
    assert w["obj_05/conv2d/kernel"].shape == (5, 5, 3, 128) and w["obj_05/dense/kernel"].shape == (32768, 128)
    assert np.all(w["obj_05/conv2d/bias"] == 0)                      # tf.layers default: zero bias, glorot-uniform kernel
    lim = np.sqrt(6.0 / (25 * 3 + 25 * 128))
    assert np.abs(w["obj_05/conv2d/kernel"]).max() <= lim
    with pytest.raises(KeyError, match="obj_05/dense/kernel"):      # like tf.train.Saver.restore: a missing variable is an error
        enc.load_weights({"conv2d_3/bias": np.ones(512, np.float32)})
    assert np.all(enc.get_weights()["obj_05/conv2d_3/bias"] == 0)   # ... and nothing was modified
    enc.load_weights({"conv2d_3/bias": np.ones(512, np.float32)}, strict=False)   # partial update; short names are accepted
    assert np.all(enc.get_weights()["obj_05/conv2d_3/bias"] == 1)
    with pytest.raises(ValueError):
        enc.load_weights({"dense/kernel": np.zeros((3, 3), np.float32)}, strict=False)
    with pytest.raises(NotImplementedError):
        Encoder(x, 128, [128], 5, [2], True)


def test_saver_round_trip(tmp_path):
    from augmentedautoencoder_b200 import _lib
    from augmentedautoencoder_b200.ae import factory
    from augmentedautoencoder_b200.ae import session as S
    from augmentedautoencoder_b200.ae.codebook import Codebook
    from augmentedautoencoder_b200.ae.encoder import Encoder
    _lib.lib()
    ds = Dataset(None, min_n_views=162, num_cyclo=12, radius=700)
    with S.variable_scope("e1"):
        x = S.placeholder(np.float32, [None, 32, 32, 3])
        enc = Encoder(x, 16, [8, 16], 5, [2, 2], False, seed=1)
        cb = Codebook(enc, ds, True)
    E = np.random.RandomState(0).randn(ds.embedding_size, 16).astype(np.float32)
    cb.embedding_normalized.assign(E)
    saver = factory.Saver([enc, cb])
    path = saver.save(None, str(tmp_path / "checkpoints" / "chkpt"), global_step=30000)
    assert path.endswith("chkpt-30000.npz")
    with S.variable_scope("e1"):
        enc2 = Encoder(S.placeholder(np.float32, [None, 32, 32, 3]), 16, [8, 16], 5, [2, 2], False, seed=2)
        cb2 = Codebook(enc2, ds, True)
    assert not np.array_equal(enc2.get_weights()["e1/conv2d/kernel"], enc.get_weights()["e1/conv2d/kernel"])
    got = factory.restore_checkpoint(None, factory.Saver([enc2, cb2]), str(tmp_path / "checkpoints"))
    assert got == path
    for k, v in enc.get_weights().items():
        assert np.array_equal(enc2.get_weights()[k], v)
    assert np.array_equal(cb2.embedding_normalized.value(), E)
    with pytest.raises(FileNotFoundError):
        factory.restore_checkpoint(None, factory.Saver([enc2]), str(tmp_path / "nothing"))
    # a checkpoint of another experiment scope must not restore "successfully" (tf.train.Saver raises NotFoundError)
    with S.variable_scope("other_scope"):
        enc3 = Encoder(S.placeholder(np.float32, [None, 32, 32, 3]), 16, [8, 16], 5, [2, 2], False, seed=3)
        cb3 = Codebook(enc3, ds, True)
    before = enc3.get_weights()["other_scope/conv2d/kernel"].copy()
    with pytest.raises(KeyError, match="other_scope/embedding_normalized|other_scope/conv2d/kernel"):
        factory.restore_checkpoint(None, factory.Saver([enc3, cb3]), str(tmp_path / "checkpoints"))
    assert np.array_equal(enc3.get_weights()["other_scope/conv2d/kernel"], before)
    with pytest.raises(KeyError, match="embedding_normalized"):     # encoder-only checkpoint into a codebook saver
        p2 = factory.Saver([enc]).save(None, str(tmp_path / "enc_only" / "chkpt"), global_step=1)
        factory.Saver([enc2, cb2]).restore(None, p2)
    factory.Saver([enc2, cb2]).restore(None, p2, strict=False)


def test_queue_pulls_one_fresh_batch_per_run():
    """ae_train.py:128 `sess.run(train_op)` dequeues a new batch every run; x and y of one run belong to the same batch."""
    import gc

    import torch
    from augmentedautoencoder_b200.ae import factory
    from augmentedautoencoder_b200.ae.session import RunContext

    class FakeSession:
        device = torch.device("cpu")

    class FakeDataset:
        shape = (4, 4, 3)

    pulls = []

    def source(n):
        i = len(pulls)
        pulls.append(i)
        return np.full((n, 4, 4, 3), i, np.float32), np.full((n, 4, 4, 3), -i, np.float32)

    q = factory.Queue(FakeDataset(), 1, 1, 2, source=source)
    seen = []
    for _ in range(6):
        ctx = RunContext(FakeSession(), {})              # what Session.run builds per call; freed (and its address reused) every time
        x, y = ctx.get(q.x), ctx.get(q.y)
        assert float(x[0, 0, 0, 0]) == -float(y[0, 0, 0, 0])     # same dequeue
        seen.append(int(x[0, 0, 0, 0]))
        del ctx
        gc.collect()
    assert seen == list(range(6)) and len(pulls) == 6


def test_tf_tensor_bundle_round_trip(tmp_path):
    from augmentedautoencoder_b200.ae.tf_checkpoint import crc32c, latest_checkpoint, read_tf_checkpoint, write_tf_checkpoint
    assert crc32c(b"123456789") == 0xE3069283                      # the standard CRC-32C check value
    rng = np.random.RandomState(0)
    tensors = {"obj/conv2d/kernel": rng.randn(5, 5, 3, 8).astype(np.float32), "obj/conv2d/bias": rng.randn(8).astype(np.float32),
               "obj/conv2d_1/kernel": rng.randn(5, 5, 8, 4).astype(np.float32), "obj/embed_obj_bbs_var": rng.randint(0, 700, (36, 4)).astype(np.int32),
               "obj/global_step": np.asarray(30000, dtype=np.int64), "obj/embedding_normalized": rng.randn(36, 16).astype(np.float32)}
    for i in range(40):                                            # > one restart interval of prefix-compressed keys
        tensors["obj/extra_%02d/v" % i] = rng.randn(3, i + 1).astype(np.float32)
    prefix = str(tmp_path / "checkpoints" / "chkpt-30000")
    write_tf_checkpoint(prefix, tensors)
    assert os.path.exists(prefix + ".index") and os.path.exists(prefix + ".data-00000-of-00001")
    got = read_tf_checkpoint(prefix, verify_crc=True)
    assert set(got) == set(tensors)
    for k, v in tensors.items():
        assert got[k].dtype == v.dtype and got[k].shape == v.shape and np.array_equal(got[k], v), k
    assert set(read_tf_checkpoint(prefix, names={"obj/conv2d/bias"})) == {"obj/conv2d/bias"}
    multi = str(tmp_path / "checkpoints" / "chkpt-30001")           # index split over several data blocks, as TensorFlow does every 4 KB
    write_tf_checkpoint(multi, tensors, entries_per_block=7)
    got = read_tf_checkpoint(multi, verify_crc=True)
    assert set(got) == set(tensors) and all(np.array_equal(got[k], v) for k, v in tensors.items())
    with open(prefix + ".index", "r+b") as f:                      # corrupt the magic -> loud failure
        f.seek(-1, 2)
        f.write(b"\x00")
    with pytest.raises(ValueError):
        read_tf_checkpoint(prefix)
    assert latest_checkpoint(str(tmp_path)) == (None, [])


def test_restore_checkpoint_reads_tensorflow_layout(tmp_path):
    from augmentedautoencoder_b200 import _lib
    from augmentedautoencoder_b200.ae import factory
    from augmentedautoencoder_b200.ae import session as S
    from augmentedautoencoder_b200.ae.codebook import Codebook
    from augmentedautoencoder_b200.ae.encoder import Encoder
    _lib.lib()
    ds = Dataset(None, min_n_views=12, num_cyclo=4, radius=700)

    def make(seed):
        with S.variable_scope("e1"):
            enc = Encoder(S.placeholder(np.float32, [None, 32, 32, 3]), 16, [8, 16], 5, [2, 2], False, seed=seed)
            return enc, Codebook(enc, ds, True)
    enc, cb = make(1)
    cb.embedding_normalized.assign(np.random.RandomState(0).randn(ds.embedding_size, 16).astype(np.float32))
    ckdir = tmp_path / "checkpoints"
    factory.Saver([enc, cb]).save_tf(None, str(ckdir / "chkpt"), global_step=20000)
    prefix = factory.Saver([enc, cb]).save_tf(None, str(ckdir / "chkpt"), global_step=30000)
    (ckdir / "checkpoint").write_text('model_checkpoint_path: "chkpt-30000"\nall_model_checkpoint_paths: "chkpt-20000"\n'
                                      'all_model_checkpoint_paths: "chkpt-30000"\n')
    enc2, cb2 = make(2)
    assert factory.restore_checkpoint(None, factory.Saver([enc2, cb2]), str(ckdir)) == prefix
    for k, v in enc.get_weights().items():
        assert np.array_equal(enc2.get_weights()[k], v)
    assert np.array_equal(cb2.embedding_normalized.value(), cb.embedding_normalized.value())
    assert factory.restore_checkpoint(None, factory.Saver([enc2, cb2]), str(ckdir), at_step=20000).endswith("chkpt-20000")
    with pytest.raises(FileNotFoundError):
        factory.restore_checkpoint(None, factory.Saver([enc2, cb2]), str(ckdir), at_step=12345)


def test_reader_on_an_independently_assembled_bundle(golden_dir):
    """SURVEY 8f N1.  tests/golden/tf_bundle/ was NOT written by this repo's writer: its index values were serialised by
    google.protobuf from TensorFlow's own message classes (TensorBoard ships them), its checksums come from TensorFlow-team
    code, and the table container was assembled from the LevelDB format description -- two data shards, several index blocks
    with shortened separators, restart points, gaps between tensors, a string entry and an unknown field
    (tests/golden/make_tf_bundle.py).  The reader must return every numeric tensor bit for bit, and restore_checkpoint must
    fill an encoder + codebook built under the same scope (auto_pose/ae/ae_factory.py:149-172)."""
    from augmentedautoencoder_b200 import _lib
    from augmentedautoencoder_b200.ae import factory
    from augmentedautoencoder_b200.ae import session as S
    from augmentedautoencoder_b200.ae.codebook import Codebook
    from augmentedautoencoder_b200.ae.encoder import Encoder
    from augmentedautoencoder_b200.ae.tf_checkpoint import latest_checkpoint, read_tf_checkpoint
    d = os.path.join(golden_dir, "tf_bundle")
    exp = np.load(os.path.join(d, "expected.npz"))
    got = read_tf_checkpoint(os.path.join(d, "chkpt-30000"), verify_crc=True)
    assert set(got) == {k.replace("|", "/") for k in exp.files}              # the DT_STRING entry is skipped, nothing else is
    for k in exp.files:
        v = got[k.replace("|", "/")]
        assert v.dtype == exp[k].dtype and v.shape == exp[k].shape and np.array_equal(v, exp[k]), k
    assert got["obj_07/global_step"].shape == () and int(got["obj_07/global_step"]) == 30000
    assert latest_checkpoint(d)[0] == os.path.join(d, "chkpt-30000")
    _lib.lib()
    ds = Dataset(None, min_n_views=12, num_cyclo=4, radius=700)
    assert ds.embedding_size == 48
    with S.variable_scope("obj_07"):
        enc = Encoder(S.placeholder(np.float32, [None, 32, 32, 3]), 16, [8, 16], 5, [2, 2], False, seed=5)
        cb = Codebook(enc, ds, True)
    assert factory.restore_checkpoint(None, factory.Saver([enc, cb]), d) == os.path.join(d, "chkpt-30000")
    w = enc.get_weights()
    for name in ("conv2d/kernel", "conv2d/bias", "conv2d_1/kernel", "conv2d_1/bias", "dense/kernel", "dense/bias"):
        assert np.array_equal(w["obj_07/" + name], exp[("obj_07/" + name).replace("/", "|")])
    assert np.array_equal(cb.embedding_normalized.value(), exp["obj_07|embedding_normalized"])
    assert np.array_equal(cb.embed_obj_bbs_var.value(), exp["obj_07|embed_obj_bbs_var"])


def test_crc32c_and_mask_agree_with_tensorflows_own_code():
    tb = pytest.importorskip("tensorboard.compat.tensorflow_stub.pywrap_tensorflow")
    from augmentedautoencoder_b200.ae.tf_checkpoint import _mask_crc, crc32c
    rng = np.random.RandomState(1)
    for n in (0, 1, 7, 64, 1000):
        b = rng.randint(0, 256, n).astype(np.uint8).tobytes()
        assert crc32c(b) == tb.crc32c(b) and _mask_crc(crc32c(b)) == tb.masked_crc32c(b)
