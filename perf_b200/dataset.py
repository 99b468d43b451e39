"""Dataset = one RGB panorama + its reference distance / normal maps.

Mirrors `/root/reference/modules/dataset/dataset.py` for the case SURVEY.md §8(f) row 3 scopes in:
``<image>_ref_distance.npy`` and ``<image>_ref_normal.npy`` exist next to the image
(`dataset.py:76-81,136-137`), so none of the 2-D predictors (Omnidata, the joint depth/normal
optimiser) is needed -- those are outside the per-ray path and raise here with that explanation.

File formats (unchanged): the image through OpenCV (BGR on disk, RGB float32 in [0,1] in memory,
`utils/utils.py:75-89`); ``*_ref_distance.npy`` [H,W] or [H,W,1] float, ``*_ref_normal.npy`` [H,W,3];
``*_ref_geometry.ply`` (binary little-endian, float xyz + uchar rgba, what ``trimesh.PointCloud.export``
writes) is re-exported after normalisation like the reference does.
"""
from __future__ import annotations

import os

import cv2 as cv
import numpy as np
import torch

from .sup_info import pano_dirs


def read_image(in_path, squeeze=True, to_torch=True, channel_first=False, factor=1):       # utils/utils.py:75-89
    img = cv.imread(in_path)
    if img is None:
        raise FileNotFoundError(in_path)
    img = img[:, :, ::-1].copy()
    if factor != 1:
        h, w, _ = img.shape
        img = cv.resize(img, (w // factor, h // factor), interpolation=cv.INTER_AREA)
    if squeeze:
        img = img.astype(np.float32) / 255.
    if to_torch:
        img = torch.from_numpy(img)
    if channel_first:
        img = img.permute(2, 0, 1)
    return img


def write_image(out_path, image):                                                            # utils/utils.py:66-72
    if torch.is_tensor(image):
        image = image.detach().cpu().numpy()
    assert (len(image.shape) == 3 and image.shape[-1] in [1, 3]) or len(image.shape) == 2
    if len(image.shape) == 3:
        image = image[:, :, ::-1].copy()
    if not cv.imwrite(out_path, image):
        raise IOError(f"cv2.imwrite failed for {out_path}")


def colorize_single_channel_image(image, color_map=cv.COLORMAP_JET):                         # utils/utils.py:92-107
    image = image.squeeze()
    assert len(image.shape) == 2
    image = (image - image.min()) / (image.max() - image.min() + 1e-6) * 255
    if torch.is_tensor(image):
        image = image.cpu().numpy()
    return cv.applyColorMap(image.astype(np.uint8), color_map)


def write_point_cloud_ply(path: str, pts: np.ndarray, colors: np.ndarray | None = None):
    """Binary little-endian PLY, float xyz (+ uchar rgba)."""
    pts = np.ascontiguousarray(pts, dtype="<f4").reshape(-1, 3)
    header = ["ply", "format binary_little_endian 1.0", f"element vertex {len(pts)}",
              "property float x", "property float y", "property float z"]
    if colors is not None:
        header += ["property uchar red", "property uchar green", "property uchar blue", "property uchar alpha"]
        rec = np.zeros(len(pts), dtype=[("p", "<f4", 3), ("c", "u1", 4)])
        c = np.asarray(colors).reshape(-1, 3)
        if c.dtype != np.uint8:
            c = (np.clip(c, 0.0, 1.0) * 255).astype(np.uint8)
        rec["p"], rec["c"][:, :3], rec["c"][:, 3] = pts, c, 255
    else:
        rec = pts
    with open(path, "wb") as f:
        f.write(("\n".join(header + ["end_header"]) + "\n").encode("ascii"))
        f.write(rec.tobytes())


class Dataset:                                                                               # dataset.py:15-128
    def __init__(self, device="cuda"):
        self.device = torch.device(device)
        self.image_path = self.ref_distance_path = self.ref_normal_path = self.ref_geometry_path = None
        self.image = self.gt_distance = self.ref_distance = self.ref_normal = None
        self.height = self.width = 0
        self.data_dir = None
        self.case_name = "wp"

    def get_joint_distance_normal(self, org_distance=None):                                  # dataset.py:69-94
        assert self.image is not None and self.height > 0 and self.width > 0
        assert self.ref_distance_path is not None and self.ref_normal_path is not None
        if not (os.path.exists(self.ref_distance_path) and os.path.exists(self.ref_normal_path)):
            raise FileNotFoundError(
                f"{self.ref_distance_path} / {self.ref_normal_path} not found: the reference would now run its monocular "
                "depth / normal predictors (PanoJointPredictor: Omnidata checkpoints), which are outside the per-ray path "
                "this library replaces.  Produce the two .npy files with the reference once, or supply your own.")
        ref_distance = torch.from_numpy(np.load(self.ref_distance_path).astype(np.float32)).to(self.device)
        ref_normal = torch.from_numpy(np.load(self.ref_normal_path).astype(np.float32)).to(self.device)
        return ref_distance, ref_normal

    def normalization(self):                                                                 # dataset.py:96-101
        scale = self.ref_distance.max().item() * 1.05
        self.ref_distance /= scale

    def save_ref_geometry(self):                                                             # dataset.py:103-119
        for path, value in ((self.ref_distance_path, self.ref_distance), (self.ref_normal_path, self.ref_normal)):
            if path is not None:                                              # atomic: other ranks may be reading it
                tmp = f"{path}.{os.getpid()}.tmp.npy"
                np.save(tmp, value.cpu().numpy())
                os.replace(tmp, path)
        pts = self.ref_point_cloud().cpu().numpy().reshape(-1, 3)
        assert self.ref_geometry_path is not None and self.ref_geometry_path[-4:] == ".ply"
        write_point_cloud_ply(self.ref_geometry_path, pts, None if self.image is None else self.image.reshape(-1, 3).cpu().numpy())

    @torch.no_grad()
    def ref_point_cloud(self):                                                               # dataset.py:121-128
        return pano_dirs(self.height, self.width, self.ref_distance.device) * self.ref_distance.squeeze()[..., None]


class WildDataset(Dataset):                                                                  # dataset.py:131-154
    def __init__(self, conf, device="cuda"):
        super().__init__(device)
        self.image_path = conf["image_path"]
        stem = ".".join(self.image_path.split(".")[:-1])
        self.ref_distance_path, self.ref_normal_path = stem + "_ref_distance.npy", stem + "_ref_normal.npy"
        self.ref_geometry_path = stem + "_ref_geometry.ply"
        self.case_name = self.image_path.split("/")[-2]
        self.image = read_image(self.image_path, to_torch=True, squeeze=True)
        resize = conf.get("image_resize")                                   # `if 'image_resize' in conf` (dataset.py:142)
        if resize is not None:
            self.width, self.height = resize
            self.image = torch.from_numpy(cv.resize(self.image.numpy(), (self.width, self.height), cv.INTER_AREA))
        else:
            self.height, self.width, _ = self.image.shape
        self.image = self.image.to(self.device)
        self.ref_distance, self.ref_normal = self.get_joint_distance_normal()
        if tuple(self.ref_distance.squeeze().shape) != (self.height, self.width):
            raise ValueError(f"{self.ref_distance_path}: shape {tuple(self.ref_distance.shape)} does not match the image "
                             f"({self.height}, {self.width})")
        self.normalization()
        from . import parallel
        if parallel.rank() == 0:                                               # under torchrun every rank builds the dataset
            self.save_ref_geometry()
