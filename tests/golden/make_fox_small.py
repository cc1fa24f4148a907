#!/usr/bin/env python
"""Builds tests/golden/fox_small/: the reference's own capture `data/fox` (BASELINE config #3; 50 hand-held 1080x1920 JPEG frames,
aabb_scale 4) reduced 6x to 180x320 so that it can travel to the GPU box as a fixture (the reference tree does not).

    python tests/golden/make_fox_small.py            # needs /root/reference; writes frames/NNNN.jpg and capture.npz

capture.npz holds what the capture's transforms_{train,test}.json hold -- per-frame file names and 4x4 camera-to-world matrices
(including the 17 frames the json lists without an image on disk, which the loader must skip, dataset.py:103-107), intrinsics scaled to
the reduced resolution, aabb_scale -- and `materialise(dst)` writes them back out in the reference's dataset layout (transforms_*.json +
images/) for NerfDataset.  Test data only."""
import json
import os
import shutil

import numpy as np

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fox_small")
SRC = "/root/reference/data/fox"
FACTOR = 6


def build():
    from PIL import Image
    os.makedirs(os.path.join(HERE, "frames"), exist_ok=True)
    out = {}
    for split in ("train", "test"):
        d = json.load(open(os.path.join(SRC, f"transforms_{split}.json")))
        names, mats, present = [], [], []
        for fr in d["frames"]:
            names.append(fr["file_path"])
            mats.append(np.array(fr["transform_matrix"], np.float64))
            p = os.path.join(SRC, fr["file_path"])
            present.append(os.path.exists(p))
            if present[-1] and split == "train":
                im = Image.open(p).convert("RGB")
                W, H = im.size
                im = im.resize((W // FACTOR, H // FACTOR), Image.LANCZOS)
                im.save(os.path.join(HERE, "frames", os.path.basename(fr["file_path"])), quality=88)
        out[f"{split}_names"] = np.array(names)
        out[f"{split}_matrices"] = np.stack(mats)
        out[f"{split}_present"] = np.array(present)
        if split == "train":
            W, H = int(d["w"]), int(d["h"])
            out["intrinsics"] = np.array([d["fl_x"] / FACTOR, d["fl_y"] / FACTOR, d["cx"] / FACTOR, d["cy"] / FACTOR, W // FACTOR, H // FACTOR], np.float64)
            out["distortion"] = np.array([d["k1"], d["k2"], d["p1"], d["p2"]], np.float64)
            out["aabb_scale"] = np.array(d["aabb_scale"])
            out["camera_angle"] = np.array([d["camera_angle_x"], d["camera_angle_y"]], np.float64)
    np.savez_compressed(os.path.join(HERE, "capture.npz"), **out)


def materialise(dst):
    """Write the fixture in the reference's dataset layout under `dst` (transforms_train.json, transforms_test.json, images/)."""
    c = np.load(os.path.join(HERE, "capture.npz"))
    os.makedirs(os.path.join(dst, "images"), exist_ok=True)
    for f in os.listdir(os.path.join(HERE, "frames")):
        shutil.copy(os.path.join(HERE, "frames", f), os.path.join(dst, "images", f))
    fl_x, fl_y, cx, cy, W, H = c["intrinsics"]
    k1, k2, p1, p2 = c["distortion"]
    for split in ("train", "test"):
        frames = [{"file_path": str(n), "transform_matrix": m.tolist()} for n, m in zip(c[f"{split}_names"], c[f"{split}_matrices"])]
        json.dump({"camera_angle_x": float(c["camera_angle"][0]), "camera_angle_y": float(c["camera_angle"][1]), "fl_x": float(fl_x), "fl_y": float(fl_y),
                   "k1": float(k1), "k2": float(k2), "p1": float(p1), "p2": float(p2), "cx": float(cx), "cy": float(cy), "w": float(W), "h": float(H),
                   "aabb_scale": int(c["aabb_scale"]), "frames": frames}, open(os.path.join(dst, f"transforms_{split}.json"), "w"))
    return dst


if __name__ == "__main__":
    build()
    print("wrote", HERE)
