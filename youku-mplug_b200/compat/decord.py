"""Stand-in for decord (dataset/video_utils/utils.py:6,12,97-117): VideoReader over a `.npy` file holding
uint8 frames [T, H, W, 3] - enough to drive the data pipeline with synthetic clips; real video decoding needs
the real decord."""
import numpy as np
import torch


class _Bridge:
    def set_bridge(self, name):
        self.name = name


bridge = _Bridge()


class VideoReader:
    def __init__(self, uri, width=None, height=None, **kwargs):
        if not (isinstance(uri, str) and uri.endswith(".npy")):
            raise RuntimeError("decord stand-in: only .npy clips [T,H,W,3] uint8 are supported (install decord for real videos)")
        self.frames = np.load(uri, mmap_mode="r")
        assert self.frames.ndim == 4 and self.frames.dtype == np.uint8

    def __len__(self):
        return self.frames.shape[0]

    def get_avg_fps(self):
        return 30.0

    def get_batch(self, indices):
        return torch.from_numpy(np.stack([np.asarray(self.frames[int(i)]) for i in indices]))
