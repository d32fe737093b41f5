"""GPU parity tests of the tensor-core path (AAE_PREC_TC_SPLIT: tcgen05 + TMA + TMEM, split-fp16 operands) against the
float64 oracle, layer by layer and end to end.  Same tolerances as the fp32 SIMT path: this path must be fp32-grade."""
import numpy as np
import pytest
import torch

from oracle import aae_oracle as O
from tests.test_gpu_parity import COS_TOL, _codebook, _enc, sess  # noqa: F401

pytestmark = pytest.mark.gpu


def test_tc_match_small_codebook_exact_fields(sess):
    """Tiny case first (one tile, B < 128): isolates descriptor / swizzle errors from pipeline errors."""
    E = O.make_codebook(3, n=64, num_cyclo=1, duplicate_cyclo_endpoints=False)
    p = O.make_encoder_params(42)
    enc = _enc(0, 8, p)
    cb = _codebook(enc, E, num_cyclo=1, max_batch=8, precision=1)
    z = np.random.RandomState(1).standard_normal((5, 128)).astype(np.float32)
    s, i = cb.match_device(torch.from_numpy(z).cuda())
    torch.cuda.synchronize()
    cos = O.cos_similarity(z.astype(np.float64), E.astype(np.float64))
    assert np.array_equal(i.cpu().numpy()[:, 0], np.argmax(cos, axis=1)), (i.cpu().numpy()[:, 0], np.argmax(cos, axis=1), s.cpu().numpy()[:, 0], cos.max(axis=1))
    assert np.max(np.abs(s.cpu().numpy()[:, 0] - cos.max(axis=1))) < 2e-6


@pytest.mark.parametrize("n_rows,batch", [(64 * 5 + 17, 100), (64 * 300, 256), (92232, 129)])
def test_tc_match_matches_oracle(sess, n_rows, batch):
    E = O.make_codebook(5, n=n_rows, num_cyclo=36 if n_rows % 36 == 0 else 1, duplicate_cyclo_endpoints=(n_rows % 36 == 0))
    p = O.make_encoder_params(42)
    enc = _enc(0, 256, p)
    cb = _codebook(enc, E, num_cyclo=36 if n_rows % 36 == 0 else 1, max_batch=256, precision=1)
    rng = np.random.RandomState(n_rows)
    z = (rng.standard_normal((batch, 128)) * rng.uniform(0.05, 50, (batch, 1))).astype(np.float32)
    s, i = cb.match_device(torch.from_numpy(z).cuda())
    s, i = s.cpu().numpy()[:, 0], i.cpu().numpy()[:, 0]
    cos64 = O.cos_similarity(z.astype(np.float64), E.astype(np.float64))
    want = np.argmax(cos64, axis=1)
    err = np.max(np.abs(s - cos64[np.arange(batch), i]))
    assert err < 2e-6, err
    for b in np.nonzero(want != i)[0]:
        assert abs(cos64[b, want[b]] - cos64[b, i[b]]) < 2e-7, (b, want[b], i[b])
    # twice in a row: the last CTA re-arms the scratch, the second launch must give the same answer
    s2, i2 = cb.match_device(torch.from_numpy(z).cuda())
    assert np.array_equal(i2.cpu().numpy()[:, 0], i) and np.array_equal(s2.cpu().numpy()[:, 0], s)


def test_tc_encoder_layers_and_latent_match_oracle(sess):
    p = O.make_encoder_params(42, bias_scale=0.05)
    enc = _enc(1, 4, p)
    xu8 = O.make_crops_u8(1234, 4)
    z = sess.run(enc.z, {enc.x: xu8})
    outs64 = O.encoder_layers(O.preprocess(xu8), p, dtype=torch.float64)
    errs = []
    for layer in range(4):
        a = enc.activation_device(layer, sess.device).cpu().numpy()
        ref = outs64[layer].numpy()
        assert a.shape == ref.shape
        errs.append(np.max(np.abs(a - ref)) / max(1.0, np.abs(ref).max()))
    z64 = outs64[5].numpy()
    errs.append(np.max(np.abs(z - z64)) / np.abs(z64).max())
    print("tc encoder relative errors per layer + latent:", ["%.2e" % e for e in errs])
    assert all(e < 1e-5 for e in errs), errs


def test_tc_end_to_end_256_crops_index_parity(sess):
    p = O.make_encoder_params(42)
    E = O.make_codebook(7)
    enc = _enc(1, 256, p)
    cb = _codebook(enc, E, max_batch=256, precision=1)
    crops = O.make_crops_u8(1234, 256)
    got = cb.nearest_rotation(sess, crops, return_idcs=True)
    with torch.cuda.device(0):
        s_dev, _ = cb.nearest_idx_device(torch.from_numpy(crops).cuda())
    want, cos = O.nearest_rotation_idcs(crops, p, E, return_cos=True)
    bad = np.nonzero(got != want)[0]
    if len(bad):
        z64 = O.encoder_forward(O.preprocess(crops[bad]), p, dtype=torch.float64)
        c64 = O.l2_normalize(z64) @ E.astype(np.float64).T
        for j, b in enumerate(bad):
            assert abs(c64[j, got[b]] - c64[j, want[b]]) < 2e-6, ("index mismatch beyond fp32 resolution", b)
    assert len(bad) <= 1
    assert np.max(np.abs(s_dev.cpu().numpy()[:, 0] - cos[np.arange(256), got])) <= COS_TOL
