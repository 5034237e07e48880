"""The turntable loop at the end of the reference's training script (entrypoints/__test_nerf.py:326-341): render every
pose of `render_poses` (dataset/dataloader.py:68-74: 160 views on a circle) with `render.render` and hand the 8-bit
frames to a writer.  The reference writes an mp4 through imageio (not in this image): the writer is a callback here
(anything with `append_data(frame)`, e.g. an imageio writer, or a list's `append`)."""
import numpy as np
import torch

from . import render as _render


def to8b(x):
    """to8b of the reference scripts: clip to [0, 1] and quantise to uint8."""
    x = x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)
    return (255 * np.clip(x, 0, 1)).astype(np.uint8)


@torch.no_grad()
def render_turntable(H, W, K, render_poses, render_kwargs, writer=None, process_group=None):
    """for pose in render_poses: rgb = render(H, W, K, c2w=pose[:3, :4], **render_kwargs_test); writer.append_data(to8b(rgb)).
    Returns the frames as one uint8 array [N, H, W, 3].  `process_group` shards every frame's rays over the ranks
    (render(..., process_group=)); every rank gets the assembled frames, only rank 0 should pass a writer."""
    frames = []
    for i in range(len(render_poses)):
        pose = torch.as_tensor(render_poses[i])[:3, :4]
        rgb, _, _, _ = _render.render(H, W, K, c2w=pose, process_group=process_group, **render_kwargs)
        frame = to8b(rgb)
        if writer is not None:
            (writer.append_data if hasattr(writer, "append_data") else writer)(np.hstack([frame]))
        frames.append(frame)
    return np.stack(frames) if frames else np.zeros((0, H, W, 3), np.uint8)
