"""Synthetic place-recognition data.

SyntheticGallery: in-memory split with the reference dataset's attribute names (q_test / db_test / test_pos,
items = (fname, pid, x, y); ibl/utils/data/dataset.py) -- images are generated from seeds by the callers.

write_synthetic_pitts_tree: a Pittsburgh-shaped tree ON DISK -- NetVLAD-style dbStruct .mat files for the
train / val / test splits and small JPEG images laid out as `raw/Pittsburgh/{images,queries}/...` -- so that
code written for the real dataset (the reference's examples/test.py) runs end to end through
`datasets.create('pitts', root, scale='30k')`, PIL decoding and the torchvision transforms."""
import os
import os.path as osp

import numpy as np


class SyntheticGallery:
    def __init__(self, root=None, n_db=1000, n_q=100, scale=None, verbose=False, seed=0):
        rng = np.random.RandomState(seed)
        self.images_dir = root
        self.db_test = [("db/%06d.jpg" % i, i, float(i), 0.0) for i in range(n_db)]
        pos = rng.randint(0, n_db, size=n_q)
        self.q_test = [("q/%06d.jpg" % i, n_db + i, float(pos[i]), 0.0) for i in range(n_q)]
        self.test_pos = [np.array([int(p)]) for p in pos]


def _scene(rng, h, w, street):
    """A smooth colour field shared by the whole `street` (places of one street look alike, as real ones do) plus
    a few rectangles of its own: compresses like a photo, similar to its neighbours, not identical."""
    img = street.copy()
    for _ in range(3):
        y0, x0 = rng.randint(0, h - 8), rng.randint(0, w - 8)
        y1, x1 = y0 + rng.randint(6, h // 3), x0 + rng.randint(6, w // 3)
        img[y0:y1, x0:x1] = rng.randint(0, 256, size=3)
    return img


def _street(rng, h, w):
    from PIL import Image
    coarse = rng.randint(0, 256, size=(6, 8, 3)).astype(np.uint8)
    return np.asarray(Image.fromarray(coarse).resize((w, h), Image.BICUBIC)).astype(np.int16)


def write_synthetic_pitts_tree(root, scale="30k", n_places=(24, 10, 40), views=2, size=(120, 160), seed=0):
    """Writes <root>/raw/pitts<scale>_{train,val,test}.mat and the JPEGs they name.

    Each split has n database places on a 100 m grid with `views` images each and one query place 4 m from
    every second database place (its image is a perturbed view of that place), so every query has exactly one
    positive place inside the 10 m / 25 m radii the reference uses (dataset.py:103-110).  Returns the root."""
    from PIL import Image
    from scipy.io import savemat
    rng = np.random.RandomState(seed)
    raw = osp.join(root, "raw")
    h, w = size
    pano = 0
    for split, n_db in zip(("train", "val", "test"), n_places):
        db_names, db_utm, q_names, q_utm = [], [], [], []
        for p in range(n_db):
            if p % 8 == 0:
                street = _street(rng, h, w)
            scene = _scene(rng, h, w, street)
            x, y = 1000.0 * (("train", "val", "test").index(split) + 1) + 100.0 * (p % 6), 100.0 * (p // 6)
            sid = "%06d" % pano
            pano += 1
            for v in range(views):
                name = osp.join("%03d" % (pano // 1000), "%s_pitch%d_yaw%d.jpg" % (sid, 1, v + 1))
                view = np.clip(np.roll(scene, 3 * v, axis=1) + rng.randint(-6, 7, size=scene.shape), 0, 255)
                _save(Image, osp.join(raw, "Pittsburgh", "images", name), view)
                db_names.append(name)
                db_utm.append((x, y))
            if p % 2 == 0:
                qsid = "%06d" % pano
                pano += 1
                name = osp.join("%03d" % (pano // 1000), "%s_pitch1_yaw1.jpg" % qsid)
                view = np.clip(np.roll(scene, -6, axis=1) + rng.randint(-40, 41, size=scene.shape), 0, 255)
                _save(Image, osp.join(raw, "Pittsburgh", "queries", name), view)
                q_names.append(name)
                q_utm.append((x + 4.0, y))
        cell = lambda names: np.array([[n] for n in names], dtype=object)
        st = {"whichSet": split, "dbImageFns": cell(db_names), "utmDb": np.asarray(db_utm, dtype=np.float64).T,
              "qImageFns": cell(q_names), "utmQ": np.asarray(q_utm, dtype=np.float64).T,
              "numImages": float(len(db_names)), "numQueries": float(len(q_names))}
        os.makedirs(raw, exist_ok=True)
        savemat(osp.join(raw, "pitts%s_%s.mat" % (scale, split)), {"dbStruct": st})
    return root


def _save(Image, path, arr):
    os.makedirs(osp.dirname(path), exist_ok=True)
    Image.fromarray(arr.astype(np.uint8)).save(path, quality=92)
