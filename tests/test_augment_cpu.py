"""Training input pipeline: the CPU restatement is pinned against OpenCV -- the library imgaug 0.4.0 delegates the
geometric ops to -- and the product's host-side tables are checked against the restatement.  No GPU needed."""
import numpy as np
import pytest

from augmentedautoencoder_b200.ae import augment as A
from oracle import augment_oracle as AO

cv2 = pytest.importorskip("cv2")

TEMPLATE_CODE = """Sequential([
    #Sometimes(0.5, PerspectiveTransform(0.05)),
    Sometimes(0.5, Affine(scale=(1.0, 1.2))),
    Sometimes(0.5, CoarseDropout( p=0.2, size_percent=0.05) ),
    Sometimes(0.5, GaussianBlur(1.2*np.random.rand())),
    Sometimes(0.5, Add((-25, 25), per_channel=0.3)),
    Sometimes(0.3, Invert(0.2, per_channel=True)),
    Sometimes(0.5, Multiply((0.6, 1.4), per_channel=0.5)),
    Sometimes(0.5, Multiply((0.6, 1.4))),
    Sometimes(0.5, ContrastNormalization((0.5, 2.2), per_channel=0.3))
    ], random_order=False)"""


def _img(seed, h=128, w=128):
    return np.random.RandomState(seed).randint(0, 256, (h, w, 3), dtype=np.uint8)


def test_warp_affine_restatement_is_bit_exact_with_opencv():
    img = _img(0)
    for s in np.linspace(1.0, 1.2, 21):
        M = AO.scale_matrix(s, 128, 128)
        ref = cv2.warpAffine(img, M, (128, 128), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
        assert np.array_equal(AO.warp_affine_u8(img, M), ref), s
    for M in (np.array([[0.9, 0.2, 5.3], [-0.15, 1.1, -3.7]]), np.array([[1.3, -0.4, 20.0], [0.25, 0.8, 11.5]])):
        ref = cv2.warpAffine(img, M, (128, 128), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
        assert np.array_equal(AO.warp_affine_u8(img, M), ref)


def test_gaussian_blur_restatement_is_bit_exact_with_opencv():
    img = _img(1, 96, 80)
    for sigma in np.concatenate([np.linspace(0.05, 1.45, 57), [0.49991074016040443, 1.2 * 0.999]]):
        assert AO.blur_ksize(sigma) == 5
        ref = cv2.GaussianBlur(img, (5, 5), sigmaX=float(sigma), sigmaY=float(sigma), borderType=cv2.BORDER_REFLECT_101)
        assert np.array_equal(AO.gaussian_blur5_u8(img, float(sigma)), ref), sigma
        assert AO.gaussian_kernel5_q8(float(sigma)).sum() == 256


def test_nearest_upsampling_map_matches_opencv():
    for dst, src in ((128, 6), (128, 4), (96, 5), (64, 6)):
        low = np.arange(src * src, dtype=np.uint8).reshape(src, src)
        ref = cv2.resize(low, (dst, dst), interpolation=cv2.INTER_NEAREST)
        m = AO.nearest_index_map(dst, src)
        assert np.array_equal(low[m][:, m], ref), (dst, src)
        assert np.array_equal(A.nearest_cells(dst, src), m)


def test_template_cfg_parses_into_the_supported_chain():
    ops = A.parse_code(TEMPLATE_CODE)
    assert [op.kind for _, op in ops] == ["Affine", "CoarseDropout", "GaussianBlur", "Add", "Invert", "Multiply", "Multiply", "ContrastNormalization"]
    assert [p for p, _ in ops] == [0.5, 0.5, 0.5, 0.5, 0.3, 0.5, 0.5, 0.5]
    aug = A.Augmenter(TEMPLATE_CODE, seed=3)
    assert aug.low == (6, 6) and 0.0 <= aug.sigma < 1.2
    with pytest.raises(NotImplementedError):
        A.parse_code("Sequential([Sometimes(0.5, PerspectiveTransform(0.05))])")


def test_host_tables_agree_with_the_restatement():
    assert np.array_equal(A.bilinear_table().astype(np.int32), AO.bilinear_table())
    for s in (1.0, 1.07, 1.2):
        M = AO.scale_matrix(s, 128, 128)
        got, want = A.affine_tables(M, 128, 128), AO.affine_fixed_point(M, 128, 128)
        for g, w in zip(got, want):
            assert np.array_equal(g.astype(np.int64), w)
    for sigma in (0.2, 0.5, 0.9, 1.19):
        assert np.array_equal(A.gaussian_taps_q8(sigma), AO.gaussian_kernel5_q8(sigma))
    aug = A.Augmenter(TEMPLATE_CODE, seed=5)
    P = aug.sample(16)
    geom, lut = aug.pack(P)
    assert geom.shape == (16, 4 + 2 * 128 + 2 * 128) and lut.shape == (16, 3, 256)
    # the composed table equals applying the value ops one after the other
    ramp = np.tile(np.arange(256, dtype=np.uint8)[None, :, None], (1, 1, 3))             # [1, 256, 3] image holding every value
    for b in range(16):
        Pb = {k: v[b:b + 1] for k, v in P.items()}
        Pb["affine_on"], Pb["drop_on"], Pb["blur_on"] = np.zeros(1, bool), np.zeros(1, bool), np.zeros(1, bool)
        want = AO.augment_batch(ramp[None], np.zeros((1, 1, 256), bool), ramp[None], Pb, 0.0, low=aug.low)[0, 0]      # [256, 3]
        assert np.array_equal(lut[b].T, want), b
    # flags and dropout bits round-trip
    for b in range(16):
        keep = (int(np.uint32(geom[b, 1])) | (int(np.uint32(geom[b, 2])) << 32))
        bits = np.array([(keep >> i) & 1 for i in range(36)], np.uint8).reshape(6, 6)
        assert np.array_equal(bits, P["drop_keep"][b])
        assert bool(geom[b, 0] & A.FLAG_AFFINE) == bool(P["affine_on"][b]) and bool(geom[b, 0] & A.FLAG_DROP) == bool(P["drop_on"][b])


def test_sampled_parameters_follow_the_cfg_distributions():
    aug = A.Augmenter(TEMPLATE_CODE, seed=11)
    P = aug.sample(4000)
    assert abs(P["affine_on"].mean() - 0.5) < 0.03 and abs(P["invert_on"].mean() - 0.3) < 0.03
    s = P["affine_M"][:, 0, 0]
    assert s.min() >= 1.0 and s.max() <= 1.2 and np.allclose(P["affine_M"][:, 0, 2], 63.5 - s * 63.5)
    assert P["add_val"].min() >= -25 and P["add_val"].max() <= 25
    same = (P["add_val"] == P["add_val"][:, :1]).all(1)
    assert abs(same.mean() - (0.7 + 0.3 / 51 ** 2)) < 0.03                      # per_channel = 0.3
    assert abs(P["drop_keep"].mean() - 0.8) < 0.01
    assert 0.6 <= P["mul1_val"].min() and P["mul1_val"].max() <= 1.4 and (P["mul2_val"] == P["mul2_val"][:, :1]).all()
    assert abs(P["invert_ch"].mean() - 0.2) < 0.02


def test_subset_chains_and_unsupported_variants():
    aug = A.Augmenter("Sequential([Sometimes(1.0, Add((10, 10))), Multiply((2.0, 2.0))])", seed=0)
    P = aug.sample(5)
    assert P["add_on"].all() and P["mul1_on"].all() and not P["affine_on"].any() and not P["blur_on"].any()
    _, lut = aug.pack(P)
    want = np.clip((np.clip(np.arange(256) + 10, 0, 255)).astype(np.float32) * np.float32(2.0), 0, 255).astype(np.uint8)
    assert np.array_equal(lut[0, 0], want) and np.array_equal(lut[4, 2], want)
    with pytest.raises(NotImplementedError):
        A.Augmenter("Sequential([Sometimes(0.5, Add((1, 2)))], random_order=True)")
    with pytest.raises(NotImplementedError):                      # value ops in another order than the kernels apply them
        A.Augmenter("Sequential([Multiply((0.5, 1.5)), Add((1, 2))])")
    with pytest.raises(NotImplementedError):
        A.Augmenter("Sequential([Sometimes(0.5, GaussianBlur(2.0))])")
    with pytest.raises(NotImplementedError):
        A.Augmenter("Sequential([Sometimes(0.5, CoarseDropout(p=0.1, size_percent=0.5))])")


def test_packed_affine_tables_match_the_restatement_per_image():
    aug = A.Augmenter(TEMPLATE_CODE, seed=9)
    P = aug.sample(32)
    P["affine_M"][3] = [[0.9, 0.2, 5.3], [-0.15, 1.1, -3.7]]          # a general matrix, not only scalings
    P["affine_on"][3] = True
    geom, _ = aug.pack(P)
    H = W = 128
    for b in range(32):
        if not P["affine_on"][b]:
            assert not geom[b, 4:].any()
            continue
        ad, bd, x0, y0 = AO.affine_fixed_point(P["affine_M"][b], H, W)
        assert np.array_equal(geom[b, 4:4 + W], ad) and np.array_equal(geom[b, 4 + W:4 + 2 * W], bd)
        assert np.array_equal(geom[b, 4 + 2 * W:4 + 2 * W + H], x0) and np.array_equal(geom[b, 4 + 2 * W + H:], y0)


@pytest.mark.parametrize("fire", [(), ("add_on",), ("invert_on",), ("contrast_on",), ("mul1_on", "mul2_on")])
def test_packed_tables_are_c_contiguous_whatever_fires(fire):
    """The kernel reads geom / lut as raw [B][...] memory: with no value op firing the table is still the stride-0 broadcast
    of the identity ramp, and a K-order astype of that view is NOT row-major (round-1 regression)."""
    aug = A.Augmenter(TEMPLATE_CODE, seed=0)
    P = aug.sample(3)
    for k in P:
        if k.endswith("_on"):
            P[k][:] = k in fire
    geom, lut = aug.pack(P)
    assert geom.flags.c_contiguous and lut.flags.c_contiguous and lut.dtype == np.uint8 and geom.dtype == np.int32
    assert lut.strides == (3 * 256, 256, 1)
    raw = np.frombuffer(lut.tobytes(order="A"), np.uint8).reshape(3, 3, 256)      # memory order, as the device sees it
    assert np.array_equal(raw, lut)
    if not fire:
        assert np.array_equal(raw, np.broadcast_to(np.arange(256, dtype=np.uint8), (3, 3, 256)))
