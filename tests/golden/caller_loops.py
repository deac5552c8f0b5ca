"""The loop bodies of the reference's callers, transcribed line for line, as functions of a namespace ``ns`` that supplies
``get_samples``, ``get_camera_from_tensor`` and a ``renderer`` -- the reference's own (make_caller_golden.py, CPU) or the
drop-ins (tests/test_gpu_callers.py, CUDA).  Nothing else differs between the two runs, which is the point of the test:
Tracker.py / Mapper.py keep working unchanged on the drop-ins.

  mapper_iteration   src/Mapper.py:498-578  (optimize_map, one joint iteration, BA on, NICE mode)
  tracker_iteration  src/Tracker.py:159-201 (optimize_cam_in_batch, RGB-D branch)
"""
import torch


def mapper_iteration(ns, c, decoders, keyframes, camera_tensor_list, cur, bound, stage, pixs_per_image, device,
                     w_color_loss=0.2):
    """keyframes: list of dicts {'depth','color','est_c2w'}; the oldest (index 0) keeps its pose (Mapper.py:375);
    cur: dict {'depth','color'} of the current frame, optimised through camera_tensor_list[-1]."""
    get_samples, get_camera_from_tensor, renderer = ns.get_samples, ns.get_camera_from_tensor, ns.renderer
    H, W, fx, fy, cx, cy = ns.H, ns.W, ns.fx, ns.fy, ns.cx, ns.cy
    optimize_frame = list(range(len(keyframes))) + [-1]
    oldest_frame = 0
    batch_rays_d_list = []
    batch_rays_o_list = []
    batch_gt_depth_list = []
    batch_gt_color_list = []

    camera_tensor_id = 0
    for frame in optimize_frame:
        if frame != -1:
            gt_depth = keyframes[frame]['depth'].to(device)
            gt_color = keyframes[frame]['color'].to(device)
            if frame != oldest_frame:
                camera_tensor = camera_tensor_list[camera_tensor_id]
                camera_tensor_id += 1
                c2w = get_camera_from_tensor(camera_tensor)
            else:
                c2w = keyframes[frame]['est_c2w']

        else:
            gt_depth = cur['depth'].to(device)
            gt_color = cur['color'].to(device)
            camera_tensor = camera_tensor_list[camera_tensor_id]
            c2w = get_camera_from_tensor(camera_tensor)

        batch_rays_o, batch_rays_d, batch_gt_depth, batch_gt_color = get_samples(
            0, H, 0, W, pixs_per_image, H, W, fx, fy, cx, cy, c2w, gt_depth, gt_color, device)
        batch_rays_o_list.append(batch_rays_o.float())
        batch_rays_d_list.append(batch_rays_d.float())
        batch_gt_depth_list.append(batch_gt_depth.float())
        batch_gt_color_list.append(batch_gt_color.float())

    batch_rays_d = torch.cat(batch_rays_d_list)
    batch_rays_o = torch.cat(batch_rays_o_list)
    batch_gt_depth = torch.cat(batch_gt_depth_list)
    batch_gt_color = torch.cat(batch_gt_color_list)

    # should pre-filter those out of bounding box depth value
    with torch.no_grad():
        det_rays_o = batch_rays_o.clone().detach().unsqueeze(-1)  # (N, 3, 1)
        det_rays_d = batch_rays_d.clone().detach().unsqueeze(-1)  # (N, 3, 1)
        t = (bound.unsqueeze(0).to(
            device)-det_rays_o)/det_rays_d
        t, _ = torch.min(torch.max(t, dim=2)[0], dim=1)
        inside_mask = t >= batch_gt_depth
    batch_rays_d = batch_rays_d[inside_mask]
    batch_rays_o = batch_rays_o[inside_mask]
    batch_gt_depth = batch_gt_depth[inside_mask]
    batch_gt_color = batch_gt_color[inside_mask]
    ret = renderer.render_batch_ray(c, decoders, batch_rays_d,
                                    batch_rays_o, device, stage,
                                    gt_depth=batch_gt_depth)
    depth, uncertainty, color = ret

    depth_mask = (batch_gt_depth > 0)

    # loss definition
    loss_rgbd = torch.abs(
        batch_gt_depth[depth_mask]-depth[depth_mask]).sum()
    if stage == 'color':
        color_loss = torch.abs(batch_gt_color - color).sum()
        weighted_color_loss = w_color_loss*color_loss
        loss_rgbd += weighted_color_loss

    loss_rgbd.backward(retain_graph=False)
    return loss_rgbd.item(), int(inside_mask.sum().item()), depth.detach(), color.detach()


def tracker_iteration(ns, c, decoders, camera_tensor, gt_depth, gt_color, bound, batch_size, Hedge, Wedge, device,
                      handle_dynamic=True, use_color_in_tracking=True, w_color_loss=0.5):
    get_samples, get_camera_from_tensor, renderer = ns.get_samples, ns.get_camera_from_tensor, ns.renderer
    H, W, fx, fy, cx, cy = ns.H, ns.W, ns.fx, ns.fy, ns.cx, ns.cy
    c2w = get_camera_from_tensor(camera_tensor)
    batch_rays_o, batch_rays_d, batch_gt_depth, batch_gt_color = get_samples(
        Hedge, H-Hedge, Wedge, W-Wedge, batch_size, H, W, fx, fy, cx, cy, c2w, gt_depth, gt_color, device)
    # should pre-filter those out of bounding box depth value
    with torch.no_grad():
        det_rays_o = batch_rays_o.clone().detach().unsqueeze(-1)  # (N, 3, 1)
        det_rays_d = batch_rays_d.clone().detach().unsqueeze(-1)  # (N, 3, 1)
        t = (bound.unsqueeze(0).to(device)-det_rays_o)/det_rays_d
        t, _ = torch.min(torch.max(t, dim=2)[0], dim=1)
        inside_mask = t >= batch_gt_depth
    batch_rays_d = batch_rays_d[inside_mask]
    batch_rays_o = batch_rays_o[inside_mask]
    batch_gt_depth = batch_gt_depth[inside_mask]
    batch_gt_color = batch_gt_color[inside_mask]

    ret = renderer.render_batch_ray(
        c, decoders, batch_rays_d, batch_rays_o,  device, stage='color',  gt_depth=batch_gt_depth)
    depth, uncertainty, color = ret

    uncertainty = uncertainty.detach()
    if handle_dynamic:
        tmp = torch.abs(batch_gt_depth-depth)/torch.sqrt(uncertainty+1e-10)
        mask = (tmp < 10*tmp.median()) & (batch_gt_depth > 0)
    else:
        mask = batch_gt_depth > 0

    # loss definition
    loss_rgbd = (torch.abs(batch_gt_depth-depth) /
                 torch.sqrt(uncertainty+1e-10))[mask].sum()

    if use_color_in_tracking:
        color_loss = torch.abs(
            batch_gt_color - color)[mask].sum()
        loss_rgbd += w_color_loss*color_loss

    loss_rgbd.backward(retain_graph=False)
    return loss_rgbd.item(), int(inside_mask.sum().item()), int(mask.sum().item())


def caller_inputs():
    """Deterministic inputs of both iterations on the tiny scene: three keyframes + the current frame."""
    import numpy as np
    import evennicer_slam_b200.synthetic as syn
    from cases import SEED
    frames = []
    for f in range(4):
        cam_t = syn.default_pose(syn.TINY_BOUND, jitter_seed=30 + f)
        depth, color, _ = syn.synthetic_frame(syn.TINY_BOUND, syn.TINY_CAM, cam_t, seed=SEED + f, zero_frac=0.1)
        depth = depth.copy()
        depth[::3, ::2] *= 1.4            # these pixels measure a depth BEHIND the scene bound: the callers' inside_mask drops them
        frames.append((cam_t.astype(np.float32), depth, color))
    return frames
