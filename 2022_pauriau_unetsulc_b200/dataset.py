"""``SulciDataset`` — point list ("bucket") -> dense binary volume + label volume, with the reference's random
rotation augmentation (reference dataset.py:33-88, 304-326).

Restated, not copied: the rotation is built with Rodrigues' formula; the order and kind of random draws is kept
(``random.uniform`` x2 for the axis, ``np.random.normal`` x1 for the angle, per training sample) so that a seeded run
visits the same augmented volumes as the reference.
"""
import math
import random

import numpy as np
import torch
from torch.utils.data import Dataset

ROT_SIGMA = math.pi / 16


def _axis_angle_matrix(axis, angle):
    ax = np.asarray(axis, dtype=np.float64)
    ax = ax / math.sqrt(float(np.dot(ax, ax)))
    K = np.array([[0.0, -ax[2], ax[1]], [ax[2], 0.0, -ax[0]], [-ax[1], ax[0], 0.0]])
    c, s = math.cos(angle), math.sin(angle)
    return c * np.eye(3) + (1.0 - c) * np.outer(ax, ax) + s * K


def random_rigid_rotation(center, sigma):
    """(R, t) of a rotation by N(0, sigma) about a uniformly drawn axis through `center`."""
    th = random.uniform(0, 2 * math.pi)
    z = random.uniform(-1, 1)
    s = np.sqrt(1 - z ** 2)
    axis = [s * np.cos(th), s * np.sin(th), z]
    R = _axis_angle_matrix(axis, np.random.normal(0, sigma))
    c = np.asarray(center, dtype=np.float64)
    return R, c - R @ c


def extract_data(graph, flip=False):
    """Point lists from a BrainVISA sulcal graph (needs ``soma.aims``; reference dataset.py:173-201)."""
    try:
        from soma import aims
    except ImportError as e:  # pragma: no cover
        raise RuntimeError("extract_data needs BrainVISA's soma.aims; pass dict_bck2/dict_names instead") from e
    to_tal = aims.GraphManip.talairach(graph)
    vs = graph['voxel_size']
    data = {'bck': [], 'nbck': [], 'bck2': [], 'vert': [], 'names': []}
    for vertex in graph.vertices():
        name = vertex['name'] if 'name' in vertex else 'unknown'
        for key in ('aims_ss', 'aims_bottom', 'aims_other'):
            if key not in vertex:
                continue
            for point in vertex[key][0].keys():
                if flip:
                    point[0] *= -1
                mm = to_tal.transform([p * v for p, v in zip(point, vs)])
                data['nbck'].append(list(point))
                data['bck'].append(list(mm))
                data['bck2'].append([int(round(mm[i] / 2)) for i in range(3)])
                data['names'].append(name)
                data['vert'].append(vertex['index'])
    return data


class SulciDataset(Dataset):
    def __init__(self, gfile_list, dict_sulci, train=True, translation_file=None, dict_bck2={}, dict_names={},
                 img_size=None, device=None, resident=True):
        """device (not in the reference): a CUDA device -> the dense volumes are built there by b2_scatter_volume
        from the (augmented) point list; the random draws and the integer point list stay on the host, so a seeded
        run visits exactly the same volumes."""
        self.device = device
        self.resident = resident     # keep each subject's point list on the device after its first use
        self.gfile_list = gfile_list
        self.dict_sulci = dict_sulci
        if 'background' not in self.dict_sulci:
            self.dict_sulci['background'] = -1
        self.train = train
        self.rot_angle = ROT_SIGMA
        self.translation_file = translation_file
        self.dict_bck2 = dict_bck2
        self.dict_names = dict_names
        self.img_size = img_size
        self._cache = {}

    def __len__(self):
        return len(self.gfile_list)

    def _points(self, gfile):
        if gfile not in self.dict_bck2:
            from soma import aims
            graph = aims.read(gfile)
            if self.translation_file is not None:
                import sigraph
                flt = sigraph.FoldLabelsTranslator()
                flt.readLabels(self.translation_file)
                flt.translate(graph)
            data = extract_data(graph)
            self.dict_bck2[gfile] = np.asarray(data['bck2'])
            self.dict_names[gfile] = np.asarray(data['names'])
        return np.asarray(self.dict_bck2[gfile]), np.asarray(self.dict_names[gfile])

    def transform(self, pts):
        if self.rot_angle is not None:
            center = (np.max(pts, axis=0) - np.min(pts, axis=0)) / 2
            R, t = random_rigid_rotation(center, self.rot_angle)
            pts = (pts @ R.T + t).astype(int)              # truncation toward zero, as the reference does
        return pts - np.min(pts, axis=0)

    # -- per-subject caches: the reference re-converts the Python lists of dict_bck2 / dict_names on every sample
    # (dataset.py:47-49, 82-86: ~5 ms per 30k-point subject); the arrays only depend on the subject
    def _cached(self, gfile):
        src = self.dict_bck2.get(gfile)
        c = self._cache.get(gfile)
        if c is None or c[0] is not src:
            pts, names = self._points(gfile)
            src = self.dict_bck2.get(gfile)
            pts = np.asarray(pts).reshape(-1, 3)
            base = pts - np.min(pts, axis=0)
            lab = np.asarray([self.dict_sulci[n] for n in names], dtype=np.int64)
            c = self._cache[gfile] = (src, base, lab, None)
        return c

    def _resident(self, gfile):
        """the subject's base points and label ids, uploaded to the device once"""
        src, base, lab, res = self._cached(gfile)
        if res is None:
            from . import ops
            res = ops.ResidentPoints(base, lab)
            self._cache[gfile] = (src, base, lab, res)
        return res

    def _draw(self, base):
        """the sample's random rigid motion (R, t) — same draws, same order as the reference — or None"""
        if not self.train or self.rot_angle is None:
            return None
        center = (np.max(base, axis=0) - np.min(base, axis=0)) / 2
        return random_rigid_rotation(center, self.rot_angle)

    def consume_draws(self, index):
        """Advances the global `random` / `np.random` streams exactly as __getitem__(index) would, without building
        the sample (data-parallel ranks skip the subjects of the other ranks but keep the seeded sequence)."""
        if self.train and self.rot_angle is not None:
            random.uniform(0, 2 * math.pi)
            random.uniform(-1, 1)
            np.random.normal(0, self.rot_angle)

    def item_size(self, index):
        """(D, H, W) of sample `index` (performs the sample's random draws), without building the volumes"""
        _, base, _, _ = self._cached(self.gfile_list[index])
        if self.img_size is not None:
            self.consume_draws(index)
            return tuple(int(v) for v in self.img_size)
        pts = self._apply(base, self._draw(base))
        return tuple(int(v) for v in np.max(pts, axis=0) + 1)

    @staticmethod
    def _apply(base, motion):
        if motion is None:
            return base
        R, t = motion
        pts = (base @ R.T + t).astype(int)                 # truncation toward zero, as the reference does
        return pts - np.min(pts, axis=0)

    def __getitem__(self, index):
        gfile = self.gfile_list[index]
        _, base, lab, _ = self._cached(gfile)
        motion = self._draw(base)
        background = self.dict_sulci['background']
        on_device = self.device is not None and torch.device(self.device).type == "cuda"
        if on_device and self.img_size is not None:
            # fixed volume size: rotation, truncation, min-shift and scatter all run on the device from the resident
            # point list (b2_scatter_volume_rot); 96 bytes of transform cross PCIe per sample
            from . import ops
            xf = None if motion is None else [float(v) for v in motion[0].reshape(-1)] + [float(v) for v in motion[1]]
            return ops.scatter_volume_rot(self._resident(gfile), xf, self.img_size, self.device,
                                          background=background, keep=self.resident)
        pts = np.array(self._apply(base, motion), dtype=int)
        size = (np.max(pts, axis=0) + 1) if self.img_size is None else self.img_size
        if on_device:
            from . import ops
            return ops.scatter_volume(pts, lab, size, self.device, background=background)
        ix = tuple(torch.as_tensor(pts[:, k], dtype=torch.long) for k in range(3))
        vol = torch.zeros(1, int(size[0]), int(size[1]), int(size[2]), dtype=torch.float)
        vol[0][ix] = 1
        labels = torch.full((int(size[0]), int(size[1]), int(size[2])), background, dtype=torch.long)
        labels[ix] = torch.as_tensor(lab, dtype=torch.long)
        return vol, labels
