"""CPU restatement (numpy) of the UNet input assembly of the event branch -- TEST INFRASTRUCTURE ONLY.

Follows /root/reference/src/event_net.py:67-87 (inference_event): permute(2,0,1) of both images, transforms.Resize(NEAREST)
when scale_factor != 1, torch.cat(dim 0), unsqueeze(0), .to(float32).  Nearest indexing is ATen's upsample_nearest2d
(what torchvision's Resize runs on tensors): src = min(int(floor(float32(dst) * float32(in / out))), in - 1).
Pinned against torchvision itself in tests/test_unet_input_cpu.py.
"""
import numpy as np


def nearest_index(n_in: int, n_out: int) -> np.ndarray:
    if n_in == n_out:
        return np.arange(n_out)
    scale = np.float32(n_in) / np.float32(n_out)
    src = np.floor(np.arange(n_out, dtype=np.float32) * scale).astype(np.int64)
    return np.minimum(src, n_in - 1)


def assemble_input(img1: np.ndarray, img2: np.ndarray, scale_factor: float = 1.0) -> np.ndarray:
    assert img1.shape == img2.shape
    H, W = img1.shape[:2]
    h, w = (int(scale_factor * H), int(scale_factor * W)) if scale_factor != 1.0 else (H, W)
    iy, ix = nearest_index(H, h), nearest_index(W, w)
    a = img1[iy][:, ix].transpose(2, 0, 1)
    b = img2[iy][:, ix].transpose(2, 0, 1)
    return np.concatenate([a, b], 0)[None].astype(np.float32)


def assemble_input_backward(g_out: np.ndarray, H: int, W: int) -> np.ndarray:
    """d loss / d img2 [H, W, 3] from d loss / d out [1, 6, h, w]."""
    h, w = g_out.shape[2], g_out.shape[3]
    iy, ix = nearest_index(H, h), nearest_index(W, w)
    g = np.zeros((H, W, 3), np.float32)
    np.add.at(g, (iy[:, None], ix[None, :]), g_out[0, 3:6].transpose(1, 2, 0))
    return g
