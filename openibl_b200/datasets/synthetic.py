"""Synthetic place-recognition split with the reference dataset's attribute names
(q_test / db_test / test_pos, items = (fname, pid, x, y); ibl/utils/data/dataset.py)."""
import numpy as np


class SyntheticGallery:
    def __init__(self, root=None, n_db=1000, n_q=100, scale=None, verbose=False, seed=0):
        rng = np.random.RandomState(seed)
        self.images_dir = root
        self.db_test = [("db/%06d.jpg" % i, i, float(i), 0.0) for i in range(n_db)]
        pos = rng.randint(0, n_db, size=n_q)
        self.q_test = [("q/%06d.jpg" % i, n_db + i, float(pos[i]), 0.0) for i in range(n_q)]
        self.test_pos = [np.array([int(p)]) for p in pos]
