"""The slice of auto_pose/ae/dataset.py the hot path touches: crop shape, the idx -> rotation table
(``viewsphere_for_embedding``, dataset.py:39-58, built on pysixd_stuff/view_sampler.py:19-188), ``embedding_size``
and the square-patch crop helper (dataset.py:354-373).  Rendering / augmentation (OpenGL, imgaug) are out of scope
(SURVEY.md section 2 rows 7, 11): ``render_embedding_image_batch`` delegates to a user-supplied renderer."""
import math

import numpy as np

from .utils import lazy_property

_GOLDEN = (1.0 + math.sqrt(5.0)) / 2.0
_ICO_VERTS = [(-1.0, _GOLDEN, 0.0), (1.0, _GOLDEN, 0.0), (-1.0, -_GOLDEN, 0.0), (1.0, -_GOLDEN, 0.0),
              (0.0, -1.0, _GOLDEN), (0.0, 1.0, _GOLDEN), (0.0, -1.0, -_GOLDEN), (0.0, 1.0, -_GOLDEN),
              (_GOLDEN, 0.0, -1.0), (_GOLDEN, 0.0, 1.0), (-_GOLDEN, 0.0, -1.0), (-_GOLDEN, 0.0, 1.0)]
_ICO_FACES = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
              (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10),
              (8, 6, 7), (9, 8, 1)]


def icosphere_points(min_n_pts, radius=1.0):
    """Hinterstoisser view sphere: subdivide an icosahedron until it has >= min_n_pts vertices, push the vertices
    to the sphere and order them ring by ring from the top pole, each ring sorted by azimuth.  The vertex numbering,
    midpoint arithmetic and ring construction reproduce view_sampler.hinter_sampling exactly (checked bit-for-bit
    against tests/golden/viewsphere_*.npz)."""
    verts = [list(v) for v in _ICO_VERTS]
    faces = list(_ICO_FACES)
    while len(verts) < min_n_pts:
        midpoint = {}
        refined = []
        for tri in faces:
            mids = []
            for a, b in ((tri[0], tri[1]), (tri[1], tri[2]), (tri[2], tri[0])):
                key = (a, b) if a < b else (b, a)
                if key not in midpoint:
                    midpoint[key] = len(verts)
                    verts.append((0.5 * (np.array(verts[key[0]]) + np.array(verts[key[1]]))).tolist())
                mids.append(midpoint[key])
            v0, v1, v2 = tri
            m01, m12, m20 = mids
            refined += [(v0, m01, m20), (m01, v1, m12), (m01, m12, m20), (m20, m12, v2)]
        faces = refined
    pts = np.array(verts)
    pts *= np.reshape(radius / np.linalg.norm(pts, axis=1), (pts.shape[0], 1))
    neighbours = {}
    for tri in faces:
        for i in range(3):
            neighbours.setdefault(tri[i], set()).update((tri[(i + 1) % 3], tri[(i + 2) % 3]))
    two_pi = 2.0 * math.pi
    azimuth = [(math.atan2(p[1], p[0]) + two_pi) % two_pi for p in pts]
    visited = [False] * len(pts)
    ring = [int(np.argmax(pts[:, 2]))]
    order = []
    while len(order) != len(pts):
        ring = sorted(ring, key=azimuth.__getitem__)
        reach = []
        for v in ring:
            order.append(v)
            visited[v] = True
            reach += [i for i in neighbours[v]]
        ring = [i for i in set(reach) if not visited[i]]  # set iteration order decides azimuth ties, as upstream
    return pts[np.array(order), :]


def look_at_rotations(pts):
    """Camera rotation for every view point: the camera looks at the origin with world +z up (OpenGL look-at), then a
    180 degree flip about x converts to the OpenCV convention (view_sampler.sample_views, view_sampler.py:160-181)."""
    c, s = math.cos(math.pi), math.sin(math.pi)
    flip = np.array([[1.0, 0.0, 0.0], [0.0, c, -s], [0.0, s, c]])
    up = np.array([0.0, 0.0, 1.0])
    out = np.empty((len(pts), 3, 3))
    for i, pt in enumerate(pts):
        fwd = -np.array(pt)
        fwd /= np.linalg.norm(fwd)
        side = np.cross(fwd, up)
        if np.count_nonzero(side) == 0:
            side = np.array([1.0, 0.0, 0.0])
        side /= np.linalg.norm(side)
        upv = np.cross(side, fwd)
        out[i] = flip.dot(np.array([[side[0], side[1], side[2]], [upv[0], upv[1], upv[2]], [-fwd[0], -fwd[1], -fwd[2]]]))
    return out


def viewsphere_rotations(min_n_views, num_cyclo, radius):
    views = look_at_rotations(icosphere_points(min_n_views, radius=radius))
    rs = np.empty((len(views) * num_cyclo, 3, 3))
    angles = np.linspace(0, 2.0 * np.pi, num_cyclo)  # both end points included -> first and last in-plane step coincide
    i = 0
    for view in views:
        for cyclo in angles:
            rot_z = np.array([[np.cos(-cyclo), -np.sin(-cyclo), 0], [np.sin(-cyclo), np.cos(-cyclo), 0], [0, 0, 1]])
            rs[i] = rot_z.dot(view)
            i += 1
    return rs


class Dataset(object):
    """Constructor signature of auto_pose/ae/dataset.py:16-36 (``Dataset(dataset_path, **kw)`` with the lower-cased
    cfg keys).  Only what the encoder / codebook path needs is kept."""

    def __init__(self, dataset_path=None, renderer=None, **kw):
        self.shape = (int(kw.get("h", 128)), int(kw.get("w", 128)), int(kw.get("c", 3)))
        self.dataset_path = dataset_path
        self._kw = dict(kw)
        self._kw.setdefault("num_cyclo", 36)
        self._kw.setdefault("min_n_views", 2562)
        self._kw.setdefault("radius", 700)
        self._renderer = renderer

    @lazy_property
    def viewsphere_for_embedding(self):
        kw = self._kw
        return viewsphere_rotations(int(kw["min_n_views"]), int(kw["num_cyclo"]), float(kw["radius"]))

    @property
    def embedding_size(self):
        return len(self.viewsphere_for_embedding)

    def render_embedding_image_batch(self, start, end):
        """(batch [n,H,W,C] float in [0,1], obj_bbs [n,4]) for codebook rows start..end (dataset.py:308-352).  Needs a
        renderer callable ``renderer(R) -> (bgr uint8 image, depth)``; OpenGL rendering is out of scope here."""
        if self._renderer is None:
            raise NotImplementedError("no renderer attached: pass renderer=callable(R)->(bgr, depth) to Dataset, or "
                                      "build the codebook with Codebook.update_embedding_from_crops")
        import cv2
        kw = self._kw
        h, w = self.shape[:2]
        pad_factor = float(kw.get("pad_factor", 1.2))
        batch = np.empty((end - start,) + self.shape)
        obj_bbs = np.empty((end - start, 4))
        for i, R in enumerate(self.viewsphere_for_embedding[start:end]):
            bgr, depth = self._renderer(R)
            ys, xs = np.nonzero(depth > 0)
            size = (depth.shape[1], depth.shape[0])
            x0, y0 = max(xs.min() - 1, 0), max(ys.min() - 1, 0)
            x1, y1 = min(xs.max() + 1, size[0] - 1), min(ys.max() + 1, size[1] - 1)
            obj_bbs[i] = [x0, y0, x1 - x0, y1 - y0]
            crop = self.extract_square_patch(bgr, obj_bbs[i], pad_factor, resize=(w, h), interpolation=cv2.INTER_NEAREST)
            batch[i] = crop / 255.
        return batch, obj_bbs

    def extract_square_patch(self, scene_img, bb_xywh, pad_factor, resize=(128, 128), interpolation=None, black_borders=False):
        """Square crop around a bbox, clipped to the image, optional blackening outside the bbox (dataset.py:354-373)."""
        import cv2
        if interpolation is None:
            interpolation = cv2.INTER_NEAREST
        x, y, w, h = np.array(bb_xywh).astype(np.int32)
        size = int(np.maximum(h, w) * pad_factor)
        left = int(np.maximum(x + w / 2 - size / 2, 0))
        right = int(np.minimum(x + w / 2 + size / 2, scene_img.shape[1]))
        top = int(np.maximum(y + h / 2 - size / 2, 0))
        bottom = int(np.minimum(y + h / 2 + size / 2, scene_img.shape[0]))
        crop = scene_img[top:bottom, left:right].copy()
        if black_borders:
            crop[:(y - top), :] = 0
            crop[(y + h - top):, :] = 0
            crop[:, :(x - left)] = 0
            crop[:, (x + w - left):] = 0
        return cv2.resize(crop, resize, interpolation=interpolation)

    # ------------------------------------------------------------------------------------------------ training batches
    def load_training_images(self, path, bg_path=None):
        """The cache the reference writes after rendering (``np.savez(current_file_name, train_x=, mask_x=, train_y=)``,
        dataset.py:101-113) and, optionally, the background image stack (``.npy``, dataset.py:229-255)."""
        data = np.load(path)
        self.train_x, self.mask_x, self.train_y = data["train_x"].astype(np.uint8), data["mask_x"], data["train_y"].astype(np.uint8)
        self.noof_training_imgs = len(self.train_x)
        if bg_path is not None:
            self.bg_imgs = np.load(bg_path).astype(np.uint8)
            self.noof_bg_imgs = len(self.bg_imgs)

    @lazy_property
    def _aug(self):
        from .augment import Augmenter
        code = self._kw.get("code")
        if code is None:
            raise NotImplementedError("no [Augmentation] CODE in the dataset arguments")
        return Augmenter(code, self.shape, seed=self._kw.get("seed"))

    def batch_device(self, batch_size, device=None):
        """Dataset.batch (dataset.py:456-495) with the image work on the GPU: draws the rendering / background indices like the
        reference, uploads the uint8 images once and returns (x, y) float32 CUDA tensors in [0, 1]."""
        import torch
        for name in ("train_x", "mask_x", "train_y", "bg_imgs"):
            if not hasattr(self, name):
                raise RuntimeError("Dataset.%s is not loaded (load_training_images / set the arrays)" % name)
        if eval(str(self._kw.get("realistic_occlusion", "False"))) or eval(str(self._kw.get("square_occlusion", "False"))):
            raise NotImplementedError("REALISTIC_OCCLUSION / SQUARE_OCCLUSION are off in the template cfg and not supported")
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        idx = np.random.choice(len(self.train_x), batch_size, replace=False)
        idx_bg = np.random.choice(len(self.bg_imgs), batch_size, replace=False)
        x = torch.from_numpy(self.train_x[idx]).to(dev, non_blocking=True)
        m = torch.from_numpy(np.ascontiguousarray(self.mask_x[idx]).astype(np.uint8)).to(dev, non_blocking=True)
        bg = torch.from_numpy(self.bg_imgs[idx_bg]).to(dev, non_blocking=True)
        y = torch.from_numpy(self.train_y[idx]).to(dev, non_blocking=True)
        xf = self._aug.augment_device(x, m, bg)
        return xf, y.to(torch.float32) / 255.0

    def batch(self, batch_size):
        """numpy (batch_x, batch_y) like the reference's ``Dataset.batch``."""
        x, y = self.batch_device(batch_size)
        return x.cpu().numpy(), y.cpu().numpy()
