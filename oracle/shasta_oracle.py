"""TEST INFRASTRUCTURE — CPU restatement of ShaSTA's affinity-estimation hot path.

This is the parity oracle, not product code: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it. The product path
(``shasta_b200``) never calls into this module and has no CPU fallback.

It restates, function by function and in the reference's own formulation (the T x D x 640/646 pair
tensors ARE materialised here, exactly as the reference does), what
``/root/reference/det3d/models/tracker/shasta.py:213-327`` computes, as plain functions over a
``{name: tensor}`` weight dict in the reference's state_dict naming. Parity pin: the reference ships no
tests or golden vectors for this path, so the oracle is pinned against outputs of the
UNMODIFIED reference run in the authoring container — ``oracle/make_golden.py`` wrote them to
``tests/golden/*.npz`` and ``tests/test_oracle.py`` checks this file against them bit for bit.

All functions take/return torch CPU tensors; ``dtype`` follows the inputs (float32 for parity,
float64 for sensitivity studies).
"""
import numpy as np
import torch
import torch.nn.functional as Fn


# --------------------------------------------------------------------------------------------------
# a0 (producer): shared_conv + NHWC permute          (shasta.py:42-47, 223-228)
# --------------------------------------------------------------------------------------------------
def shared_conv_nhwc(w, x, eps=1e-5):
    """``self.shared_conv(x).permute(0,2,3,1).contiguous()`` in inference mode: Conv2d(512->64, 3x3, padding 1,
    bias) -> BatchNorm2d with running statistics -> ReLU. ``w`` holds the reference's state_dict entries
    ``shared_conv.0.{weight,bias}``, ``shared_conv.1.{weight,bias,running_mean,running_var}``; x is (N,512,H,W)."""
    y = Fn.conv2d(x, w["shared_conv.0.weight"], w["shared_conv.0.bias"], padding=1)
    y = Fn.batch_norm(y, w["shared_conv.1.running_mean"], w["shared_conv.1.running_var"], w["shared_conv.1.weight"],
                      w["shared_conv.1.bias"], training=False, eps=eps)
    return torch.relu(y).permute(0, 2, 3, 1).contiguous()


# --------------------------------------------------------------------------------------------------
# a1: box -> 5 sample points          (shasta.py:121-161, box_torch_ops.py:24-59,145-158,184-203)
# --------------------------------------------------------------------------------------------------
def corners_nd_2d(dims):
    """box_torch_ops.py:24-59 for ndim=2, origin=0.5: clockwise corners from the minimum point."""
    norm = torch.tensor([[-0.5, -0.5], [-0.5, 0.5], [0.5, 0.5], [0.5, -0.5]], dtype=dims.dtype)
    return dims.view(-1, 1, 2) * norm.view(1, 4, 2)


def rotation_2d(points, angles):
    """box_torch_ops.py:145-158: x' = x cos + y sin ; y' = -x sin + y cos (einsum aij,jka->aik)."""
    rot_sin = torch.sin(angles)
    rot_cos = torch.cos(angles)
    rot_mat_T = torch.stack([torch.stack([rot_cos, -rot_sin]), torch.stack([rot_sin, rot_cos])])
    return torch.einsum("aij,jka->aik", (points, rot_mat_T))


def center_to_corner_box2d(centers, dims, angles):
    """box_torch_ops.py:184-203."""
    corners = corners_nd_2d(dims)
    corners = rotation_2d(corners, angles)
    corners = corners + centers.view(-1, 1, 2)
    return corners


def box_points(box):
    """shasta.py:143-159 (num_point == 5). box (M,7) -> (5M,3), point-major blocks
    [centres; fronts; backs; lefts; rights]."""
    center2d = box[:, :2]
    height = box[:, 2:3]
    dim2d = box[:, 3:5]
    rotation_y = box[:, -1]
    corners = center_to_corner_box2d(center2d, dim2d, rotation_y)
    front = torch.cat([(corners[:, 0] + corners[:, 1]) / 2, height], dim=-1)
    back = torch.cat([(corners[:, 2] + corners[:, 3]) / 2, height], dim=-1)
    left = torch.cat([(corners[:, 0] + corners[:, 3]) / 2, height], dim=-1)
    right = torch.cat([(corners[:, 1] + corners[:, 2]) / 2, height], dim=-1)
    return torch.cat([box[:, :3], front, back, left, right], dim=0)


# --------------------------------------------------------------------------------------------------
# a2: metric -> pixel, clamped bilinear gather, 5-point regroup
#                                      (bird_eye_view.py:18-41, center_utils.py:92-121)
# --------------------------------------------------------------------------------------------------
def absl_to_relative(points, pc_start, voxel_size, out_stride):
    """bird_eye_view.py:18-22 — subtraction then two successive divisions, in that order."""
    a1 = (points[..., 0] - pc_start[0]) / voxel_size[0] / out_stride
    a2 = (points[..., 1] - pc_start[1]) / voxel_size[1] / out_stride
    return a1, a2


def bilinear_interpolate(im, x, y):
    """center_utils.py:92-121. im (H,W,C); weights use the CLAMPED integer coordinates, so a point outside
    the map hits the same pixel twice with weights +a and -a (result ~0, not the border value)."""
    x0 = torch.floor(x).long()
    x1 = x0 + 1
    y0 = torch.floor(y).long()
    y1 = y0 + 1
    x0 = torch.clamp(x0, 0, im.shape[1] - 1)
    x1 = torch.clamp(x1, 0, im.shape[1] - 1)
    y0 = torch.clamp(y0, 0, im.shape[0] - 1)
    y1 = torch.clamp(y1, 0, im.shape[0] - 1)
    Ia = im[y0, x0]
    Ib = im[y1, x0]
    Ic = im[y0, x1]
    Id = im[y1, x1]
    wa = (x1.type_as(x) - x) * (y1.type_as(y) - y)
    wb = (x1.type_as(x) - x) * (y - y0.type_as(y))
    wc = (x - x0.type_as(x)) * (y1.type_as(y) - y)
    wd = (x - x0.type_as(x)) * (y - y0.type_as(y))
    return Ia * wa[:, None] + Ib * wb[:, None] + Ic * wc[:, None] + Id * wd[:, None]


def gather_box_features(bev, boxes7, pc_start, voxel_size, out_stride, num_point=5):
    """shasta.py:231-233 + bird_eye_view.py:24-41. bev (B,H,W,C), boxes7 (B,M,7) -> (B,M,num_point*C)."""
    out = []
    for b in range(bev.shape[0]):
        pts = box_points(boxes7[b])
        xs, ys = absl_to_relative(pts, pc_start, voxel_size, out_stride)
        fm = bilinear_interpolate(bev[b], xs, ys)
        sec = fm.shape[0] // num_point
        out.append(torch.cat([fm[i * sec:(i + 1) * sec] for i in range(num_point)], dim=1))
    return torch.stack(out)


# --------------------------------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------------------------------
def _linear(w, name, x):
    return Fn.linear(x, w[name + ".weight"], w[name + ".bias"])


def _mlp(w, prefix, idxs, x):
    """nn.Sequential of Linear/ReLU pairs with Linear at the listed indices, no ReLU after the last."""
    for n, i in enumerate(idxs):
        x = _linear(w, "%s.%d" % (prefix, i), x)
        if n + 1 < len(idxs):
            x = torch.relu(x)
    return x


# --------------------------------------------------------------------------------------------------
# a3/a4: anchors                                       (shasta.py:49-57,69-76,241-247,260-267)
# --------------------------------------------------------------------------------------------------
def anchor_shapes(w, feature, prev_feature):
    """shasta.py:241-244. -> newborn, fp (from current feature), dead, fn (from previous), each (B,1,F)."""
    B = feature.shape[0]
    flat = feature.reshape(B, -1)
    pflat = prev_feature.reshape(B, -1)
    srcs = [flat, flat, pflat, pflat]
    return [torch.abs(_mlp(w, "aug_shape.%d" % i, (0, 2), srcs[i])).reshape(B, 1, -1) for i in range(4)]


def anchor_boxes(w, det7, prev7):
    """shasta.py:260-267 — inputs are the boxes BEFORE back-projection; dims [3:6] are abs'd."""
    B = det7.shape[0]
    srcs = [det7.reshape(B, -1), det7.reshape(B, -1), prev7.reshape(B, -1), prev7.reshape(B, -1)]
    out = []
    for i in range(4):
        a = _mlp(w, "aug_dets.%d" % i, (0, 2), srcs[i]).reshape(B, 1, -1)
        out.append(torch.cat((a[:, :, :3], torch.abs(a[:, :, 3:6]), a[:, :, 6:]), dim=-1))
    return out  # newborn, fp, dead_trk, fn


# --------------------------------------------------------------------------------------------------
# a5: hand-designed residuals                                          (shasta.py:277-283)
# --------------------------------------------------------------------------------------------------
def handcrafted_residual(prev_aug, det_aug, num_feats):
    eps = 1e-10
    dist = ((prev_aug[:, :, :num_feats].unsqueeze(2) - det_aug[:, :, :num_feats].unsqueeze(1)) ** 2).sum(dim=-1)
    dist = Fn.normalize(dist)  # p=2 over dim=1 (the T axis), eps 1e-12
    dim = torch.abs(torch.log(prev_aug[:, :, 3:6].unsqueeze(2) + eps)
                    - torch.log(det_aug[:, :, 3:6].unsqueeze(1) + eps)).sum(dim=-1)
    dist = dist + dim
    rot = torch.sqrt((torch.cos(prev_aug[:, :, 6].unsqueeze(2)) - torch.cos(det_aug[:, :, 6].unsqueeze(1))) ** 2
                     + (torch.sin(prev_aug[:, :, 6].unsqueeze(2)) - torch.sin(det_aug[:, :, 6].unsqueeze(1))) ** 2)
    return dist + rot


# --------------------------------------------------------------------------------------------------
# full forward                                                         (shasta.py:213-327)
# --------------------------------------------------------------------------------------------------
def forward(w, bev, prev_bev, det_boxes, prev_det_boxes, num_feats=3, pc_start=(-54, -54),
            voxel_size=(0.075, 0.075), out_stride=8, mutate=True, return_intermediates=False):
    """Reference forward from the 64-channel NHWC maps. ``det_boxes`` (B,M,11) is back-projected in
    place when ``mutate`` (shasta.py:270 writes through the view). Returns (matched1, matched2[, inter])."""
    if not mutate:
        det_boxes = det_boxes.clone()
    prev7 = prev_det_boxes[:, :, :7]
    det7 = det_boxes[:, :, :7]
    vel = det_boxes[:, :, 7:9]
    dt = det_boxes[:, :, 9].unsqueeze(-1)

    feature = gather_box_features(bev, det7, pc_start, voxel_size, out_stride)
    prev_feature = gather_box_features(prev_bev, prev7, pc_start, voxel_size, out_stride)
    inter = {"feature": feature, "prev_feature": prev_feature}

    newborn_g, fp_g, dead_g, fn_g = anchor_shapes(w, feature, prev_feature)
    feature = torch.cat((feature, dead_g, fn_g), dim=1)            # D axis: dets + dead + FN
    prev_feature = torch.cat((prev_feature, newborn_g, fp_g), dim=1)  # T axis: prev + newborn + FP

    newborn, fp, dead_trk, fn = anchor_boxes(w, det7, prev7)
    inter.update(newborn=newborn, fp=fp, dead_trk=dead_trk, fn=fn,
                 aug_shape=[newborn_g, fp_g, dead_g, fn_g])

    det7[:, :, :2] = det7[:, :, :2] - vel * dt                       # in place, through the view

    prev_aug = torch.cat((prev7, newborn, fp), dim=1)
    det_aug = torch.cat((det7, dead_trk, fn), dim=1)
    B, T, D = prev_aug.shape[0], prev_aug.shape[1], det_aug.shape[1]
    F = feature.shape[-1]

    residual_dist = handcrafted_residual(prev_aug, det_aug, num_feats)

    pf = prev_feature.unsqueeze(2).expand(B, T, D, F)
    cf = feature.unsqueeze(1).expand(B, T, D, F)
    fused_shape = torch.cat([pf, cf], dim=3).reshape(B, T * D, 2 * F)
    residual_shape = _mlp(w, "fuse_shape", (0, 2, 4, 6), fused_shape).view(B, T, D)

    pb = prev_aug[:, :, :num_feats].unsqueeze(2).expand(B, T, D, num_feats)
    cb = det_aug[:, :, :num_feats].unsqueeze(1).expand(B, T, D, num_feats)
    fused_boxes = torch.cat([pb, cb], dim=3).reshape(B, T * D, 2 * num_feats)
    residual_fused = _mlp(w, "fuse_det", (0, 2, 4), fused_boxes).view(B, T, D)

    fused_all = torch.cat([pf, pb, cf, cb], dim=-1).reshape(B, T * D, -1)
    coeff = _mlp(w, "res_coeff", (0, 2, 4), fused_all).view(B, T, D, 3)
    alpha, beta, omega = coeff[..., 0], coeff[..., 1], coeff[..., 2]

    residual = alpha * residual_fused + beta * residual_dist + omega * residual_shape
    matched = _mlp(w, "aff", (0, 2, 4, 6, 8, 10), residual)
    matched1 = torch.softmax(matched[:, :-2, :], dim=2)
    matched2 = torch.softmax(matched[:, :, :-2], dim=1)
    if return_intermediates:
        inter.update(residual=residual, logits=matched, residual_dist=residual_dist,
                     residual_shape=residual_shape, residual_fused=residual_fused, coeff=coeff,
                     prev_aug=prev_aug, det_aug=det_aug)
        return matched1, matched2, inter
    return matched1, matched2


# --------------------------------------------------------------------------------------------------
# a12: training loss                                     (tools/nusc_shasta/train.py:201-211)
# --------------------------------------------------------------------------------------------------
def affinity_loss(matched1, matched2, gt):
    eps = 1e-10
    gt1 = gt[:, :-2, :]
    gt2 = gt[:, :, :-2]
    lf = (gt1 * -torch.log(matched1 + eps)).sum()
    if gt1.sum() > 0:
        lf = lf / gt1.sum()
    lb = (gt2 * -torch.log(matched2 + eps)).sum()
    if gt2.sum() > 0:
        lb = lb / gt2.sum()
    return (lf + lb) / 2.0


# --------------------------------------------------------------------------------------------------
# a13: consumer decode                               (tools/nusc_shasta/eval.py:126-181)
# --------------------------------------------------------------------------------------------------
def decode(matched1, matched2, n_prev, n_det):
    """Pure-function restatement of the eval loop for ONE frame pair (matched1 (M,M+2), matched2 (M+2,M)).
    Returns dict(dead, fn, fn_score, keep_prev, keep_dets, newborn, det_score) with python lists, in the
    iteration order of the reference."""
    m1 = matched1.detach().cpu().numpy()
    m2 = matched2.detach().cpu().numpy()
    dead, fn, fn_score, keep_prev = [], [], [], []
    if n_prev > 0:
        A = np.concatenate((m1[:n_prev, :n_det], m1[:n_prev, -2:]), axis=1)
        k_idx = A.argmax(axis=1)
        vals = A.max(axis=1)
        for n in range(n_prev):
            val, k = float(vals[n]), int(k_idx[n])
            if val > 0.5 and k == A.shape[1] - 2:
                dead.append(n)
                continue
            if val > 0.5 and k == A.shape[1] - 1:
                fn.append(n)
                fn_score.append(1.0 - float(A[n, -2]))
                continue
            keep_prev.append(n)
        Bm = np.concatenate((m2[keep_prev, :n_det], m2[-2:, :n_det]), axis=0)
    else:
        Bm = m2[-2:, :n_det]
    keep_dets, newborn, det_score = [], [], []
    if n_det > 0:
        n_idx = Bm.argmax(axis=0)
        vals = Bm.max(axis=0)
        for k in range(n_det):
            val, n = float(vals[k]), int(n_idx[k])
            if val > 0.7 and n == Bm.shape[0] - 1:
                continue
            newborn.append(bool(val > 0.5 and n == Bm.shape[0] - 2))
            det_score.append(1.0 - float(Bm[-1, k]))
            keep_dets.append(k)
    return dict(dead=dead, fn=fn, fn_score=fn_score, keep_prev=keep_prev, keep_dets=keep_dets,
                newborn=newborn, det_score=det_score, row_argmax=(k_idx.tolist() if n_prev > 0 else []),
                col_argmax=(n_idx.tolist() if n_det > 0 else []))


def weights_to_torch(wnp, dtype=torch.float32):
    return {k: torch.from_numpy(np.ascontiguousarray(v)).to(dtype) for k, v in wnp.items()}
