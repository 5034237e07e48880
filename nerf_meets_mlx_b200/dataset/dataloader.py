"""Mirror of mlx_nerf/dataset/dataloader.py: the Blender synthetic-scene on-disk format
(`transforms_{train,val,test}.json` + RGBA PNGs).  Pure host code; PNGs are decoded with PIL (the reference uses
imageio.v2.imread, which is not in this image -- both return the same uint8 [H, W, 4] array)."""
import json
import os

import numpy as np
import torch
from PIL import Image

from ..ops import pose


def _imread(fname):
    with Image.open(fname) as im:
        return np.asarray(im)


def load_blender_data(basedir, half_res: bool = False, testskip=1):
    """load_blender_data (dataset/dataloader.py:20-93) -> (imgs [N,H,W,4] f32 in [0,1], poses [N,4,4] f32,
    render_poses [160,4,4] f32 torch, [H, W, focal], i_split).

    Declared deviation: `half_res=True` cannot run in the reference (`Image.fromarray` of a float32 RGBA array raises,
    :85); here each frame is resized from its uint8 pixels with the same LANCZOS filter and then normalised."""
    splits = ["train", "val", "test"]
    metas = {}
    for s in splits:
        with open(os.path.join(basedir, f"transforms_{s}.json"), "r") as fp:
            metas[s] = json.load(fp)

    all_imgs, all_poses, counts = [], [], [0]
    for s in splits:
        meta = metas[s]
        skip = 1 if (s == "train" or testskip == 0) else testskip  # every training frame, every testskip-th otherwise
        imgs, poses = [], []
        for frame in meta["frames"][::skip]:
            imgs.append(_imread(os.path.join(basedir, frame["file_path"] + ".png")))
            poses.append(np.array(frame["transform_matrix"]))
        imgs = np.array(imgs)
        if half_res:
            imgs = np.array([np.asarray(Image.fromarray(im).resize((im.shape[1] // 2, im.shape[0] // 2), Image.Resampling.LANCZOS))
                             for im in imgs])
        all_imgs.append((imgs / 255.0).astype(np.float32))  # all 4 channels kept
        all_poses.append(np.array(poses).astype(np.float32))
        counts.append(counts[-1] + all_imgs[-1].shape[0])

    i_split = [np.arange(counts[i], counts[i + 1]) for i in range(len(splits))]
    imgs = np.concatenate(all_imgs, 0)
    poses = np.concatenate(all_poses, 0)
    H, W = imgs[0].shape[:2]  # already halved with half_res: focal below is then full-res focal / 2 (:80)
    camera_angle_x = float(meta["camera_angle_x"])  # of the last split read, as in the reference (:64)
    focal_length = 0.5 * W / np.tan(0.5 * camera_angle_x)
    render_poses = torch.stack([pose.pose_spherical(theta=angle, phi=-30.0, radius=4.0)
                                for angle in np.linspace(-180, 180, 160 + 1)[:-1]], dim=0)
    return imgs, poses, render_poses, [H, W, focal_length], i_split


def post_load_blender_data(i_split, images, is_white_bkgd):
    """post_load_blender_data (dataset/dataloader.py:96-113): near/far = 2/6, RGBA -> RGB (composited on white when
    `is_white_bkgd`)."""
    i_train, i_val, i_test = i_split
    near, far = 2.0, 6.0
    if is_white_bkgd:
        images = images[..., :3] * images[..., -1:] + (1.0 - images[..., -1:])
    else:
        images = images[..., :3]
    return i_train, i_val, i_test, near, far, images
