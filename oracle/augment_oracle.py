"""CPU restatement of the training input pipeline of auto_pose/ae/dataset.py:456-495 (``Dataset.batch``): background paste by
mask, then the imgaug chain of the training cfg (auto_pose/ae/cfg/train_template.cfg:26-37).  TEST INFRASTRUCTURE ONLY -- imported
by tests/ (and nothing under augmentedautoencoder_b200/).

What the reference executes is imgaug 0.4.0 (aae_py37_tf26.yml:125), which is NOT installable here.  Pinning status per op:

* Affine(scale)            imgaug builds a 2x3 matrix and calls cv2.warpAffine(INTER_LINEAR, BORDER_CONSTANT, 0): ``warp_affine_u8`` is
                           checked bit-for-bit against cv2.warpAffine (tests/test_augment_cpu.py).  The matrix itself
                           (scale about (w/2 - 0.5, h/2 - 0.5)) follows imgaug's source as remembered: parity unpinned.
* CoarseDropout            low-resolution Binomial mask, nearest-neighbour upsampling through cv2.resize(INTER_NEAREST): the
                           index map is checked against cv2.resize.  Mask sampling is random by nature.
* GaussianBlur(sigma)      for uint8 imgaug 0.4.0 calls cv2.GaussianBlur(ksize 5 for sigma < 1.5, BORDER_REFLECT_101): ``gaussian_blur5_u8``
                           restates OpenCV's bit-exact fixed-point path and is checked bit-for-bit against cv2.GaussianBlur.
* Add / Invert             integer look-up tables: exact by construction.
* Multiply / ContrastNormalization   look-up tables built in float32 and truncated to uint8 as imgaug 0.4.0's uint8 paths do
                           (as remembered: parity unpinned; differences could only be +-1 on table entries).
"""
import numpy as np

# ----------------------------------------------------------------------------------------------------------- warpAffine
_AB_BITS, _INTER_BITS = 10, 5


def bilinear_table():
    """OpenCV initInterTab2D(INTER_LINEAR, fixpt): 32 x 32 sub-pixel positions, four int16 weights summing to 32768."""
    tab = np.zeros((32 * 32, 4), np.int32)
    t = np.arange(32, dtype=np.float32) / np.float32(32)
    c = np.stack([np.float32(1) - t, t], 1).astype(np.float32)
    for i in range(32):
        for j in range(32):
            w = np.array([c[i, 0] * c[j, 0], c[i, 0] * c[j, 1], c[i, 1] * c[j, 0], c[i, 1] * c[j, 1]], np.float32)
            it = np.rint(w * np.float32(32768)).astype(np.int32)
            diff = int(it.sum()) - 32768
            if diff != 0:
                mk, big = 0, 0
                for k in range(4):
                    if it[k] < it[mk]:
                        mk = k
                    elif it[k] > it[big]:
                        big = k
                if diff < 0:
                    it[big] -= diff
                else:
                    it[mk] -= diff
            tab[i * 32 + j] = it
    return tab


_TAB = None


def affine_fixed_point(M, h, w):
    """Fixed-point source coordinates of cv2.warpAffine for the forward matrix M [2,3]: per column (adelta, bdelta) and per
    row (X0, Y0), all int64, such that  X = (X0[y] + adelta[x]) >> 5,  Y = (Y0[y] + bdelta[x]) >> 5  (5 fractional bits)."""
    M = np.array(M, np.float64).reshape(2, 3).copy()
    D = M[0, 0] * M[1, 1] - M[0, 1] * M[1, 0]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = M[1, 1] * D, M[0, 0] * D
    M[0, 0] = A11
    M[0, 1] *= -D
    M[1, 0] *= -D
    M[1, 1] = A22
    b1 = -M[0, 0] * M[0, 2] - M[0, 1] * M[1, 2]
    b2 = -M[1, 0] * M[0, 2] - M[1, 1] * M[1, 2]
    M[0, 2], M[1, 2] = b1, b2
    xs = np.arange(w, dtype=np.float64)
    ys = np.arange(h, dtype=np.float64)
    scale = float(1 << _AB_BITS)
    adelta = np.rint(M[0, 0] * xs * scale).astype(np.int64)
    bdelta = np.rint(M[1, 0] * xs * scale).astype(np.int64)
    rd = (1 << _AB_BITS) // 32 // 2
    X0 = np.rint((M[0, 1] * ys + M[0, 2]) * scale).astype(np.int64) + rd
    Y0 = np.rint((M[1, 1] * ys + M[1, 2]) * scale).astype(np.int64) + rd
    return adelta, bdelta, X0, Y0


def warp_affine_u8(src, M):
    """cv2.warpAffine(src, M, (w, h), flags=INTER_LINEAR, borderMode=BORDER_CONSTANT, borderValue=0) for uint8 [h,w,c]."""
    global _TAB
    if _TAB is None:
        _TAB = bilinear_table()
    h, w = src.shape[:2]
    adelta, bdelta, X0, Y0 = affine_fixed_point(M, h, w)
    srcp = src.astype(np.int64)
    X = (X0[:, None] + adelta[None, :]) >> (_AB_BITS - _INTER_BITS)
    Y = (Y0[:, None] + bdelta[None, :]) >> (_AB_BITS - _INTER_BITS)
    sx, sy = X >> _INTER_BITS, Y >> _INTER_BITS
    wts = _TAB[(Y & 31) * 32 + (X & 31)]                         # [h,w,4]
    acc = np.zeros(src.shape, np.int64)
    for k, (dy, dx) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
        yy, xx = sy + dy, sx + dx
        ok = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)
        v = np.where(ok[..., None], srcp[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)], 0)
        acc += v * wts[..., k:k + 1]
    return np.clip((acc + (1 << 14)) >> 15, 0, 255).astype(np.uint8)


def scale_matrix(s, h, w):
    """imgaug 0.4.0 Affine(scale=s): scaling about the image centre (w/2 - 0.5, h/2 - 0.5) (``shift_add=(0.5, 0.5)``)."""
    cx, cy = w / 2.0 - 0.5, h / 2.0 - 0.5
    return np.array([[s, 0.0, cx - s * cx], [0.0, s, cy - s * cy]], np.float64)


# ----------------------------------------------------------------------------------------------------------- Gaussian blur
def gaussian_kernel5_q8(sigma):
    """The 5-tap kernel cv2.GaussianBlur uses for uint8 images (getGaussianKernelFixedPoint_ED): exp(-x^2 / 2 sigma^2) normalised
    in double, the two outer taps rounded to 8 fractional bits with error diffusion (the rounding error of tap 0 is carried
    into tap 1), the centre = 256 - 2 * (tap0 + tap1) so that the taps sum to exactly 1.0."""
    x = np.arange(5, dtype=np.float64) - 2.0
    k = np.exp(-x * x / (2.0 * sigma * sigma))
    k /= k.sum()
    kq = np.zeros(5, np.int64)
    err = 0.0
    for i in range(2):
        adj = k[i] * 256.0 + err
        v0 = int(np.rint(adj))
        err = adj - v0
        kq[i] = kq[4 - i] = v0
    kq[2] = 256 - 2 * int(kq[0] + kq[1])
    return kq


def reflect101(i, n):
    i = np.where(i < 0, -i, i)
    return np.where(i >= n, 2 * n - 2 - i, i)


def gaussian_blur5_u8(img, sigma):
    """cv2.GaussianBlur(img, (5, 5), sigmaX=sigma, sigmaY=sigma, borderType=BORDER_REFLECT_101) for uint8 [h,w,c]."""
    kq = gaussian_kernel5_q8(sigma)
    h, w = img.shape[:2]
    p = img.astype(np.int64)
    cols = [reflect101(np.arange(w) + d, w) for d in range(-2, 3)]
    rows = [reflect101(np.arange(h) + d, h) for d in range(-2, 3)]
    hsum = sum(kq[i] * p[:, cols[i]] for i in range(5))         # 8.8 fixed point
    vsum = sum(kq[i] * hsum[rows[i]] for i in range(5))         # 8.16
    return np.clip((vsum + (1 << 15)) >> 16, 0, 255).astype(np.uint8)


def blur_ksize(sigma):
    """imgaug 0.4.0 blur_gaussian_: kernel size from sigma (always 5 for the template's sigma < 1.2)."""
    if sigma < 3.0:
        k = 3.3 * sigma
    elif sigma < 5.0:
        k = 2.9 * sigma
    else:
        k = 2.6 * sigma
    k = int(max(k, 5))
    return k + 1 if k % 2 == 0 else k


# ----------------------------------------------------------------------------------------------------------- nearest upsample
def nearest_index_map(dst, src):
    """cv2.resize(INTER_NEAREST): source index of every destination index (double arithmetic as in OpenCV)."""
    ifx = 1.0 / (float(dst) / float(src))
    return np.minimum(np.floor(np.arange(dst, dtype=np.float64) * ifx).astype(np.int64), src - 1)


# ----------------------------------------------------------------------------------------------------------- look-up tables
def lut_add(value):
    return np.clip(np.arange(256, dtype=np.int16) + int(value), 0, 255).astype(np.uint8)


def lut_multiply(m):
    t = np.arange(256, dtype=np.float32) * np.float32(m)
    return np.clip(t, 0, 255).astype(np.uint8)


def lut_contrast(alpha):
    t = np.float32(127) + np.float32(alpha) * (np.arange(256, dtype=np.float32) - np.float32(127))
    return np.clip(t, 0, 255).astype(np.uint8)


# ----------------------------------------------------------------------------------------------------------- whole pipeline
def augment_batch(x, mask, bg, params, sigma, low=(6, 6)):
    """x, bg: uint8 [B,H,W,C]; mask: bool [B,H,W] (True = background pixel); params: dict of per-image arrays as produced by
    augmentedautoencoder_b200.ae.augment.Augmenter.sample (see there for the fields).  Returns uint8 [B,H,W,C]."""
    B, H, W, C = x.shape
    out = np.empty_like(x)
    rmap, cmap = nearest_index_map(H, low[0]), nearest_index_map(W, low[1])
    for b in range(B):
        img = x[b].copy()
        img[mask[b]] = bg[b][mask[b]]                                         # dataset.py:473
        if params["affine_on"][b]:
            img = warp_affine_u8(img, params["affine_M"][b])
        if params["drop_on"][b]:
            keep = params["drop_keep"][b].astype(bool)                       # [lh, lw], True = keep
            img = img * keep[rmap][:, cmap][..., None].astype(np.uint8)
        if params["blur_on"][b] and sigma > 1e-3:
            img = gaussian_blur5_u8(img, sigma)
        if params["add_on"][b]:
            for c in range(C):
                img[..., c] = lut_add(params["add_val"][b, c])[img[..., c]]
        for c in range(C):
            if params["invert_on"][b] and params["invert_ch"][b, c]:
                img[..., c] = 255 - img[..., c]
        if params["mul1_on"][b]:
            for c in range(C):
                img[..., c] = lut_multiply(params["mul1_val"][b, c])[img[..., c]]
        if params["mul2_on"][b]:
            for c in range(C):
                img[..., c] = lut_multiply(params["mul2_val"][b, c])[img[..., c]]
        if params["contrast_on"][b]:
            for c in range(C):
                img[..., c] = lut_contrast(params["contrast_val"][b, c])[img[..., c]]
        out[b] = img
    return out
