"""Seeded synthetic frame pairs and weights of nuScenes shape.

Everything is produced by a counter-based integer hash (splitmix64), so a (seed, shape) pair yields the
same numbers on every machine and library version — the golden fixtures under ``tests/golden`` store
only outputs plus an input checksum and regenerate the inputs from here.

Box layout follows the reference dataset (det3d/datasets/nuscenes/nuscenes.py:230-232,273-275):
``[x, y, z, w, l, h, yaw, vx, vy, dt, score]``, float32, leading ``n_real`` rows real and the rest exactly 0
(nuscenes.py:207,249).
"""
import math

import numpy as np

_MASK = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _MASK
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _MASK
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _MASK
    return z ^ (z >> np.uint64(31))


def _name_seed(name):
    h = np.uint64(1469598103934665603)
    with np.errstate(over="ignore"):
        for ch in name.encode():
            h = ((h ^ np.uint64(ch)) * np.uint64(1099511628211)) & _MASK
    return h


def hash_uniform(n, seed, stream=0):
    """n float64 values in [0,1) with 24 random bits each (exactly representable in fp32)."""
    with np.errstate(over="ignore"):
        base = _splitmix64(np.uint64(seed) * np.uint64(0x100000001B3) + np.uint64(stream))
        idx = np.arange(n, dtype=np.uint64)
        bits = _splitmix64(idx ^ base)
    return (bits >> np.uint64(40)).astype(np.float64) / float(1 << 24)


def hash_normal(n, seed, stream=0):
    u1 = hash_uniform(n, seed, stream * 2 + 1000003)
    u2 = hash_uniform(n, seed, stream * 2 + 1000004)
    return np.sqrt(-2.0 * np.log(1.0 - u1)) * np.cos(2.0 * math.pi * u2)


def head_param_shapes(max_obj, num_feats=3, share_conv_channel=64, num_point=5):
    """Ordered {name: shape} of the head's trainable tensors (reference shasta.py:49-106; names are the
    state_dict contract). shared_conv is outside the timed path and not listed."""
    F = share_conv_channel * num_point
    M = max_obj
    shapes = {}

    def lin(name, out_f, in_f):
        shapes[name + ".weight"] = (out_f, in_f)
        shapes[name + ".bias"] = (out_f,)

    for i in range(4):
        lin("aug_shape.%d.0" % i, (M * F) // 64, M * F)
        lin("aug_shape.%d.2" % i, F, (M * F) // 64)
    lin("fuse_shape.0", F // 8, 2 * F)
    lin("fuse_shape.2", F // 16, F // 8)
    lin("fuse_shape.4", F // 32, F // 16)
    lin("fuse_shape.6", 1, F // 32)
    for i in range(4):
        lin("aug_dets.%d.0" % i, (7 * M) // 32, 7 * M)
        lin("aug_dets.%d.2" % i, 7, (7 * M) // 32)
    lin("fuse_det.0", 32, 2 * num_feats)
    lin("fuse_det.2", 8, 32)
    lin("fuse_det.4", 1, 8)
    lin("res_coeff.0", 32 + F // 8, 2 * num_feats + 2 * F)
    lin("res_coeff.2", 8 + F // 32, 32 + F // 8)
    lin("res_coeff.4", 3, 8 + F // 32)
    lin("aff.0", 128, M + 2)
    lin("aff.2", 64, 128)
    lin("aff.4", 32, 64)
    lin("aff.6", 64, 32)
    lin("aff.8", 128, 64)
    lin("aff.10", M + 2, 128)
    return shapes


def make_weights(max_obj, seed=0, num_feats=3, peaky=0.0):
    """Deterministic nn.Linear-style init: U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias
    (same distribution family as torch's default, reproducible without torch's RNG).
    ``peaky`` > 0 multiplies ``aff.10`` by that gain so some affinities cross the 0.5/0.7 decode
    thresholds."""
    out = {}
    for name, shape in head_param_shapes(max_obj, num_feats).items():
        layer = name.rsplit(".", 1)[0]
        fan_in = head_param_shapes(max_obj, num_feats)[layer + ".weight"][1]
        bound = 1.0 / math.sqrt(fan_in) if fan_in > 0 else 0.0
        n = int(np.prod(shape))
        u = hash_uniform(n, seed, int(_name_seed(name) % np.uint64(1 << 31)))
        w = ((2.0 * u - 1.0) * bound).astype(np.float32).reshape(shape)
        if peaky and layer == "aff.10":
            w = (w * np.float32(peaky)).astype(np.float32)
        out[name] = w
    return out


def make_frame_pairs(batch, max_obj, height, width, seed, pc_start=(-54.0, -54.0), pixel=0.6,
                     channels=64, n_real=None, with_maps=True):
    """One batch of synthetic frame pairs.

    Returns dict(det_boxes (B,M,11), prev_det_boxes (B,M,11), bev (B,H,W,C), prev_bev (B,H,W,C),
    n_det (B,), n_prev (B,)) as float32 numpy arrays. Real rows: x,y ~ U over the map extent with 2 %
    pushed just outside it (exercises the clamped border taps, center_utils.py:103-106); z ~ N(-1,1);
    w,l,h ~ U(0.5,4.5); yaw ~ U(-pi,pi); v ~ N(0,3); dt = 0.5; score ~ U(0.1,1). 70 % of the current
    boxes are the previous box moved by v*dt plus N(0,0.3) noise, the rest are fresh.
    """
    B, M = batch, max_obj
    ext_x = width * pixel
    ext_y = height * pixel
    x0, y0 = pc_start
    s = 0

    def U(n):
        nonlocal s
        s += 1
        return hash_uniform(n, seed, s)

    def N(n):
        nonlocal s
        s += 1
        return hash_normal(n, seed, s)

    def fresh(n):
        b = np.zeros((n, 11), np.float64)
        b[:, 0] = x0 + U(n) * ext_x
        b[:, 1] = y0 + U(n) * ext_y
        out = U(n) < 0.02
        side = U(n) < 0.5
        push = U(n) * 7.2 + 1e-3
        b[:, 0] = np.where(out & side, x0 + ext_x + push, np.where(out & ~side, x0 - push, b[:, 0]))
        out_y = U(n) < 0.02
        push_y = U(n) * 7.2 + 1e-3
        b[:, 1] = np.where(out_y, y0 + ext_y + push_y, b[:, 1])
        b[:, 2] = -1.0 + N(n)
        b[:, 3:6] = 0.5 + 4.0 * U(3 * n).reshape(n, 3)
        b[:, 6] = (2.0 * U(n) - 1.0) * math.pi
        b[:, 7:9] = 3.0 * N(2 * n).reshape(n, 2)
        b[:, 9] = 0.5
        b[:, 10] = 0.1 + 0.9 * U(n)
        return b

    prev = np.zeros((B, M, 11), np.float64)
    cur = np.zeros((B, M, 11), np.float64)
    n_prev = np.zeros(B, np.int64)
    n_det = np.zeros(B, np.int64)
    lo = max(1, M // 2)
    for b in range(B):
        if n_real is None:
            n_prev[b] = lo + int(U(1)[0] * (M - lo + 1))
            n_det[b] = lo + int(U(1)[0] * (M - lo + 1))
        else:
            n_prev[b], n_det[b] = n_real
        n_prev[b] = min(n_prev[b], M)
        n_det[b] = min(n_det[b], M)
        p = fresh(int(n_prev[b]))
        prev[b, : n_prev[b]] = p
        c = fresh(int(n_det[b]))
        k = min(int(n_prev[b]), int(n_det[b]))
        moved = p[:k].copy()
        moved[:, 0:2] += moved[:, 7:9] * moved[:, 9:10] + 0.3 * N(2 * k).reshape(k, 2)
        moved[:, 6] += 0.05 * N(k)
        moved[:, 10] = c[:k, 10]
        keep = U(k) < 0.7
        c[:k] = np.where(keep[:, None], moved, c[:k])
        cur[b, : n_det[b]] = c
    res = dict(det_boxes=cur.astype(np.float32), prev_det_boxes=prev.astype(np.float32),
               n_det=n_det, n_prev=n_prev)
    if with_maps:
        n = B * height * width * channels
        res["bev"] = np.maximum(hash_normal(n, seed, 7001), 0.0).astype(np.float32).reshape(
            B, height, width, channels)
        res["prev_bev"] = np.maximum(hash_normal(n, seed, 7002), 0.0).astype(np.float32).reshape(
            B, height, width, channels)
    return res


def checksum(*arrays):
    """Order-sensitive 64-bit checksum of float32 arrays' bit patterns (fixture input guard)."""
    h = np.uint64(0)
    with np.errstate(over="ignore"):
        for a in arrays:
            bits = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64).ravel()
            idx = np.arange(bits.size, dtype=np.uint64)
            h = _splitmix64(h ^ np.bitwise_xor.reduce(_splitmix64(bits + (idx << np.uint64(32)))))
    return int(h)
