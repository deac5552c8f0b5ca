"""CPU oracle for the callers' RGB-D loss glue (numpy restatement).  TEST INFRASTRUCTURE ONLY.

Restates /root/reference/src/Mapper.py:553-562 and /root/reference/src/Tracker.py:180-196 with their gradients with
respect to the renderer outputs (depth float64, colour float32); ``torch.median`` is the lower middle element.
Parity pinning: ``tests/test_rgbd_loss_cpu.py`` checks value and gradients against those lines run with torch autograd.
"""
from __future__ import annotations

import numpy as np


def mapper_loss(gt_depth, gt_color, depth, color, w_color=0.2, use_color=True):
    gd = gt_depth.astype(np.float64)
    diff = gd - depth.astype(np.float64)
    m = gd > 0
    loss = np.abs(diff[m]).sum()
    g_depth = np.where(m, -np.sign(diff), 0.0)
    g_color = np.zeros(color.shape, np.float32)
    if use_color:
        dc = gt_color.astype(np.float64) - color.astype(np.float64)
        loss += w_color * np.abs(dc).sum()
        g_color = (-w_color * np.sign(dc)).astype(np.float32)
    return loss, g_depth, g_color


def tracker_loss(gt_depth, gt_color, depth, uncertainty, color, w_color=0.2, use_color=True, handle_dynamic=True):
    gd = gt_depth.astype(np.float64)
    diff = gd - depth.astype(np.float64)
    den = np.sqrt(uncertainty.astype(np.float64) + 1e-10)
    tmp = np.abs(diff) / den
    m = gd > 0
    if handle_dynamic:
        med = np.sort(tmp)[(tmp.size - 1) // 2]                 # torch.median: lower of the two middle values
        m = (tmp < 10 * med) & m
    loss = tmp[m].sum()
    g_depth = np.where(m, -np.sign(diff) / den, 0.0)
    g_color = np.zeros(color.shape, np.float32)
    if use_color:
        dc = gt_color.astype(np.float64) - color.astype(np.float64)
        loss += w_color * np.abs(dc)[m].sum()
        g_color = np.where(m[:, None], -w_color * np.sign(dc), 0.0).astype(np.float32)
    return loss, g_depth, g_color
