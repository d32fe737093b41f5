"""CPU tests that pin the oracle: golden vectors produced by the reference's own host code
(tests/golden/make_golden.py) plus structural known-answer tests for the TF-semantics restatement."""
import os

import numpy as np
import pytest
import torch

from oracle import aae_oracle as O


def g(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


# ------------------------------------------------------------------ pinned against reference code
def test_viewsphere_small_matches_reference(golden_dir):
    ref = g(golden_dir, "viewsphere_162x12.npz")["R"]
    got = O.viewsphere_for_embedding(162, 12)
    assert got.shape == ref.shape == (162 * 12, 3, 3)
    assert np.array_equal(got, ref)


def test_viewsphere_full_matches_reference(golden_dir):
    ref = g(golden_dir, "viewsphere_2562x36.npz")
    got = O.viewsphere_for_embedding(2562, 36)
    assert got.shape == (92232, 3, 3)
    assert np.array_equal(got[::36], ref["view_R"])
    assert np.array_equal(got[:72], ref["first_rows"])
    assert np.array_equal(got[-36:], ref["last_rows"])
    assert np.array_equal(got[ref["probe_idx"]], ref["probe_R"])
    assert np.allclose([got.sum(), np.abs(got).sum()], ref["checksum"], rtol=0, atol=1e-6)
    # F8: np.linspace includes both end points -> cyclo 0 and cyclo 35 are the same rotation
    assert np.max(np.abs(got[0] - got[35])) < 1e-15


def test_preprocess_matches_reference_feed(golden_dir):
    s = g(golden_dir, "select.npz")
    assert str(s["fed_dtype"]) == "float64"  # the reference divides in float64, TF casts on feed
    assert np.array_equal(O.preprocess(s["x_probe"][None])[0], s["fed_probe"])
    lut = s["u8_over_255_f32"]
    assert np.array_equal((np.arange(256, dtype=np.float32) / np.float32(255.0)), lut)  # F2: fp32 divide is exact
    assert np.count_nonzero(np.arange(256, dtype=np.float32) * np.float32(1 / 255.0) != lut) > 0


def _golden_cos(n=92232):
    rng = np.random.RandomState(5)
    cos_b = rng.standard_normal((6, n)).astype(np.float32)
    cos_b[1, 777] = cos_b[1, 40000] = cos_b[1].max() + 1.0
    cos_b[2, 36 * 100 + 35] = cos_b[2, 36 * 100] = cos_b[2].max() + 2.0
    cos_b[3, 5] = cos_b[3].max() + 3.0
    return cos_b


def test_index_selection_matches_reference(golden_dir):
    s = g(golden_dir, "select.npz")
    cos_b = _golden_cos()
    assert np.array_equal(O.select_indices(cos_b), s["idc_plain"])
    assert O.select_indices(cos_b)[1] == 777 and O.select_indices(cos_b)[2] == 3600
    assert np.array_equal(O.select_indices(cos_b, upright=True), s["idc_upright"])
    top8 = O.select_indices(cos_b[:1], top_n=8)
    assert np.array_equal(top8, s["idc_top8"])
    rs = O.viewsphere_for_embedding()
    assert np.array_equal(rs[s["idc_plain"]], s["r_batch"])
    assert np.array_equal(rs[s["idc_plain"]], s["r_nn_batch"])
    assert np.array_equal(rs[s["idc_plain"][:1]].squeeze(), s["r_single"])


def test_pose_lift_matches_reference(golden_dir):
    p = g(golden_dir, "pose_lift.npz")
    rs = O.viewsphere_for_embedding()
    for i in range(3):
        idcs = np.atleast_1d(p[f"c{i}_idcs"])
        bbs = {int(k): v for k, v in zip(idcs, p[f"c{i}_bbs_at_idcs"])}
        table = np.zeros((92232, 4), dtype=np.int32)
        for k, v in bbs.items():
            table[k] = v
        r, t = O.auto_pose6d_lift(idcs, rs, table, p[f"c{i}_bb"], p["k_test"], p["k_train"], float(p["radius"]))
        assert np.array_equal(r, p[f"c{i}_Rs"]) and np.array_equal(t, p[f"c{i}_ts"])
    table = np.zeros((92232, 4), dtype=np.int32)
    table[int(p["depth_idx"][0])] = p["depth_bb"]
    r, t = O.auto_pose6d_lift(p["depth_idx"], rs, table, p["c2_bb"], p["k_test"], p["k_train"], float(p["radius"]), depth_pred=812.5)
    assert np.array_equal(r, p["depth_Rs"]) and np.array_equal(t, p["depth_ts"])


# ------------------------------------------------------------------ TF-semantics known answers
def test_same_padding_is_asymmetric_for_stride2():
    assert O._same_pads(128, 5, 2) == (1, 2)  # F4
    assert O._same_pads(8, 5, 1) == (2, 2)
    # a one-hot kernel tap reads pixel (2*o + kh - 1): with tap (0,0), output (1,1) must see input (1,1)
    x = np.zeros((1, 8, 8, 1), np.float32)
    x[0, 1, 1, 0] = 1.0
    k = np.zeros((5, 5, 1, 1), np.float32)
    k[0, 0, 0, 0] = 1.0
    y = O.conv2d_same(torch.from_numpy(x), torch.from_numpy(k), torch.zeros(1), 2, None).numpy()
    assert y[0, 1, 1, 0] == 1.0 and y.sum() == 1.0


@pytest.mark.parametrize("stride,hw,ci,co", [(2, 8, 3, 4), (1, 6, 2, 3), (2, 16, 5, 2)])
def test_conv_matches_independent_loops(stride, hw, ci, co):
    rng = np.random.RandomState(0)
    x = rng.rand(2, hw, hw, ci).astype(np.float32)
    k = rng.randn(5, 5, ci, co).astype(np.float32)
    b = rng.randn(co).astype(np.float32)
    want = O.conv2d_same_loops(x, k, b, stride)
    got = O.conv2d_same(torch.from_numpy(x).double(), torch.from_numpy(k).double(), torch.from_numpy(b).double(), stride, None).numpy()
    assert np.allclose(got, want, rtol=0, atol=1e-12)
    got32 = O.conv2d_same(torch.from_numpy(x), torch.from_numpy(k), torch.from_numpy(b), stride, None).numpy()
    assert np.allclose(got32, want, rtol=1e-5, atol=1e-5)


def test_flatten_is_hwc_order():
    p = O.make_encoder_params(1, num_filters=(4, 8), in_hw=16, strides=(2, 2), latent=6)
    x = np.random.RandomState(1).rand(2, 16, 16, 3).astype(np.float32)
    outs = O.encoder_layers(x, p, strides=(2, 2))
    conv_last, flat, z = outs[-3], outs[-2], outs[-1]
    assert flat.shape == (2, 4 * 4 * 8)
    assert flat[1, (2 * 4 + 3) * 8 + 5] == conv_last[1, 2, 3, 5]
    assert np.allclose(z.numpy(), flat.numpy() @ p["dense/kernel"] + p["dense/bias"], atol=1e-5)


def test_l2_normalize_epsilon_and_unit_norm():
    z = np.random.RandomState(2).randn(5, 128).astype(np.float32)
    z[3] = 0
    q = O.l2_normalize(z)
    assert np.allclose(np.linalg.norm(q[[0, 1, 2, 4]], axis=1), 1, atol=1e-6)
    assert np.all(q[3] == 0)  # 0 * rsqrt(1e-12) = 0


def test_f32_vs_f64_encoder_and_argmax_small():
    p = O.make_encoder_params(3, num_filters=(8, 16), in_hw=32, strides=(2, 2), latent=16)
    x = O.preprocess(O.make_crops_u8(4, 8, hw=32))
    z32 = O.encoder_forward(x, p, strides=(2, 2))
    z64 = O.encoder_forward(x, p, strides=(2, 2), dtype=torch.float64)
    assert np.max(np.abs(z32 - z64)) < 1e-5
    cb = O.make_codebook(7, n=360, j=16, num_cyclo=36)
    assert np.array_equal(cb[35], cb[0])
    c32, c64 = O.cos_similarity(z32, cb), O.cos_similarity(z64, cb.astype(np.float64))
    assert np.max(np.abs(c32 - c64)) < 1e-6


def test_decoder_shapes_bootstrap_and_grads_small():
    enc = O.make_encoder_params(5, num_filters=(4, 8), in_hw=16, strides=(2, 2), latent=8, bias_scale=0.1)
    dec = O.make_decoder_params(6, num_filters=(4, 8), out_hw=16, strides=(2, 2), latent=8, bias_scale=0.1, n_encoder_convs=2)
    assert set(dec) == {"dense_1/kernel", "dense_1/bias", "conv2d_2/kernel", "conv2d_2/bias", "conv2d_3/kernel", "conv2d_3/bias"}
    x = np.random.RandomState(3).rand(3, 16, 16, 3).astype(np.float32)
    y = np.random.RandomState(4).rand(3, 16, 16, 3).astype(np.float32)
    loss, rec, grads = O.ae_forward_loss(x, y, enc, dec, with_grads=True)
    assert rec.shape == (3, 16, 16, 3) and 0 < rec.min() and rec.max() < 1
    l2 = (y.reshape(3, -1) - rec.reshape(3, -1)) ** 2
    k = l2.shape[1] // 4
    want = np.sort(l2, axis=1)[:, -k:].mean()
    assert abs(loss - want) < 1e-6
    loss64, _, grads64 = O.ae_forward_loss(x, y, enc, dec, dtype=torch.float64, with_grads=True)
    assert abs(loss - loss64) < 1e-6
    for kk in grads:
        assert np.allclose(grads[kk], grads64[kk], atol=1e-5)


def test_resize_nearest_is_pixel_duplication():
    x = torch.arange(2 * 3 * 3 * 1, dtype=torch.float32).reshape(2, 3, 3, 1)
    y = O.resize_nearest_2x(x, (6, 6))
    for i in range(6):
        for j in range(6):
            assert torch.equal(y[:, i, j], x[:, i // 2, j // 2])


def test_tf_adam_differs_from_torch_adam_eps_placement():
    rng = np.random.RandomState(0)
    p = rng.randn(64).astype(np.float32)
    g_ = (1e-6 * rng.randn(64)).astype(np.float32)  # tiny grads make the eps placement visible
    m = np.zeros_like(p)
    v = np.zeros_like(p)
    p1, m1, v1 = O.tf_adam_step(p, g_, m, v, 1)
    tp = torch.nn.Parameter(torch.from_numpy(p.copy()))
    opt = torch.optim.Adam([tp], lr=2e-4, eps=1e-8)
    tp.grad = torch.from_numpy(g_)
    opt.step()
    # same m/v recursion; different epsilon placement -> different step when sqrt(v) ~ eps
    assert np.allclose(m1, 0.1 * g_) and np.allclose(v1, 0.001 * g_ * g_)
    assert not np.allclose(p1, tp.detach().numpy(), rtol=0, atol=1e-9)
    lr_t = 2e-4 * np.sqrt(1 - 0.999) / (1 - 0.9)
    assert np.allclose(p1, p - lr_t * m1 / (np.sqrt(v1) + 1e-8), atol=1e-9)
