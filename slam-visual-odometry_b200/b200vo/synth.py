"""Synthetic KITTI-shaped sequences with known ground-truth poses.

Nothing like this exists in the reference (its data/ directory is untracked); it is the input
generator for tests and bench.py, not part of the hot path.  A textured corridor (ground, ceiling,
two side walls, boards at staggered depths) is rendered by per-pixel ray/plane intersection with the intrinsics the reference
synthesises from the image size, fx = fy = W, cx = W/2, cy = H/2
(/root/reference/src/visual_odometry.cpp:90-98).  torch is used only as an array library so the
same code renders on the CPU (tests, golden fixtures) and on the GPU (bench).
"""
import math

import numpy as np
import torch


def make_texture(size=1024, seed=1234):
    """Tileable multi-scale random-rectangle texture, u8 (size,size)."""
    rng = np.random.default_rng(seed)
    tex = np.full((size, size), 128.0, np.float32)
    for (count, lo, hi, amp) in ((60, 96, 256, 70), (400, 24, 96, 60), (3000, 6, 24, 55), (6000, 3, 8, 35)):
        xs = rng.integers(0, size, count)
        ys = rng.integers(0, size, count)
        ws = rng.integers(lo, hi, count)
        hs = rng.integers(lo, hi, count)
        vs = rng.uniform(-amp, amp, count)
        for x, y, w, h, v in zip(xs, ys, ws, hs, vs):
            yy = (np.arange(y, y + h) % size)[:, None]
            xx = (np.arange(x, x + w) % size)[None, :]
            tex[yy, xx] = 128.0 + v + 0.35 * (tex[yy, xx] - 128.0)
    tex += rng.normal(0, 2.0, tex.shape).astype(np.float32)
    return np.clip(tex, 0, 255).astype(np.uint8)


def camera_matrix(width, height):
    return np.array([[float(width), 0, width / 2.0], [0, float(width), height / 2.0], [0, 0, 1.0]])


def make_trajectory(nframes, seed=5678, step=1.0):
    """Smooth forward motion with gentle yaw/pitch and sway.  Returns R_wc (n,3,3), C_w (n,3)."""
    rng = np.random.default_rng(seed)
    ph = rng.uniform(0, 2 * math.pi, 4)
    R, C = [], []
    z = 0.0
    for i in range(nframes):
        yaw = 0.10 * math.sin(0.045 * i + ph[0]) + 0.04 * math.sin(0.21 * i + ph[1])
        pitch = 0.015 * math.sin(0.13 * i + ph[2])
        x = 0.45 * math.sin(0.03 * i + ph[3])
        y = 0.05 * math.sin(0.17 * i + ph[0])
        z += step * (1.0 + 0.2 * math.sin(0.07 * i + ph[1]))
        cy, sy = math.cos(yaw), math.sin(yaw)
        cp, sp = math.cos(pitch), math.sin(pitch)
        Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
        Rx = np.array([[1, 0, 0], [0, cp, -sp], [0, sp, cp]])
        R.append(Ry @ Rx)
        C.append(np.array([x, y, z]))
    return np.stack(R), np.stack(C)


def relative_pose_gt(R_wc, C_w, i):
    """(R_c2_c1, unit t_c2_c1) between frame i and i+1: x_c2 = R x_c1 + t."""
    R1, R2 = R_wc[i], R_wc[i + 1]
    R = R2.T @ R1
    t = R2.T @ (C_w[i] - C_w[i + 1])
    return R, t / np.linalg.norm(t)


# Scene (camera: x right, y down, z forward).  Four textured planes form a corridor narrow enough that
# structure starts ~5 m ahead -- (normal axis, offset, texture axes) -- and three families of
# fronto-parallel boards repeat along z at staggered periods: (z0, period, x range, y range).  The boards
# give the depth discontinuities and near-range parallax that make the translation direction of a
# monocular pair observable to a degree or two (round-1 scene: everything >= 9 m away, 10-20 deg).
_PLANES = ((1, 1.5, (0, 2)), (1, -2.0, (0, 2)), (0, -2.5, (2, 1)), (0, 3.0, (2, 1)))
_BOARDS = ((2.0, 7.0, (0.9, 1.9), (-0.9, 1.5)), (5.5, 7.0, (-1.9, -0.9), (-0.6, 1.5)), (4.0, 11.0, (-0.5, 0.5), (-2.0, -0.7)))
_BOARDS_AHEAD = 6          # boards of each family tested per ray (visible range ~ 40 m)
_BOARD_NEAR = 2.0          # the nearest board is at least this far ahead (closer ones only blur the frame)
_DETAIL_GAIN, _DETAIL_FREQ, _DETAIL_RANGE_M = 0.7, 4.3, 9.0     # fine texture layer, faded out with distance (a crude LOD)


@torch.no_grad()
def render_frames(width, height, R_wc, C_w, texture, device="cpu", texels_per_m=40.0, chunk=16):
    """Render frames -> uint8 tensor (n, height, width) on `device`."""
    dev = torch.device(device)
    n = R_wc.shape[0]
    tex = torch.as_tensor(texture, device=dev).float()
    ts = tex.shape[0]
    fx = float(width)
    detail_range = _DETAIL_RANGE_M * fx / 1241.0        # the detail layer aliases beyond ~fx/260 m
    u = (torch.arange(width, device=dev, dtype=torch.float32) - width / 2.0) / fx
    v = (torch.arange(height, device=dev, dtype=torch.float32) - height / 2.0) / fx
    dc = torch.stack([u[None, :].expand(height, width), v[:, None].expand(height, width),
                      torch.ones(height, width, device=dev)], -1)          # (H,W,3)
    out = torch.empty((n, height, width), dtype=torch.uint8, device=dev)
    Rt = torch.as_tensor(R_wc, device=dev, dtype=torch.float32)
    Ct = torch.as_tensor(C_w, device=dev, dtype=torch.float32)

    def sample(tu, tv):                                                     # bilinear, tiled
        u0 = torch.floor(tu); v0 = torch.floor(tv)
        fu = tu - u0; fv = tv - v0
        iu0 = torch.remainder(u0, ts).long(); iv0 = torch.remainder(v0, ts).long()
        iu1 = (iu0 + 1) % ts; iv1 = (iv0 + 1) % ts
        return (tex[iv0, iu0] * (1 - fu) * (1 - fv) + tex[iv0, iu1] * fu * (1 - fv)
                + tex[iv1, iu0] * (1 - fu) * fv + tex[iv1, iu1] * fu * fv)

    for s in range(0, n, chunk):
        R = Rt[s:s + chunk]
        C = Ct[s:s + chunk]
        dw = torch.einsum("nij,hwj->nhwi", R, dc)                          # (b,H,W,3)
        # pass 1: the nearest surface of every ray and its texture coordinates (no texture access yet)
        best_t = torch.full(dw.shape[:3], 1e9, device=dev)
        b_tu = torch.zeros_like(best_t); b_tv = torch.zeros_like(best_t)
        b_du = torch.zeros_like(best_t); b_dv = torch.zeros_like(best_t)

        def take(ok, t, tu, tv, du, dv):
            nonlocal best_t, b_tu, b_tv, b_du, b_dv
            best_t = torch.where(ok, t, best_t)
            b_tu = torch.where(ok, tu, b_tu); b_tv = torch.where(ok, tv, b_tv)
            b_du = torch.where(ok, du, b_du); b_dv = torch.where(ok, dv, b_dv)

        for pi, (ax, off, (ta, tb)) in enumerate(_PLANES):
            denom = dw[..., ax]
            t = (off - C[:, ax])[:, None, None] / torch.where(denom.abs() < 1e-9, torch.full_like(denom, 1e-9), denom)
            hit = C[:, None, None, :] + t[..., None] * dw
            take((t > 0.05) & (t < best_t), t, hit[..., ta] * texels_per_m + 137.0 * pi, hit[..., tb] * texels_per_m + 71.0 * pi,
                 torch.full_like(t, 31.0 * pi), torch.full_like(t, 17.0 * pi))
        dz = dw[..., 2].clamp(min=1e-3)
        for bi, (z0, per, (xa, xb), (ya, yb)) in enumerate(_BOARDS):
            k0 = torch.ceil((C[:, 2] + _BOARD_NEAR - z0) / per)
            for j in range(_BOARDS_AHEAD):
                k = k0 + j
                t = (z0 + k * per - C[:, 2])[:, None, None] / dz
                hx = C[:, 0, None, None] + t * dw[..., 0]
                hy = C[:, 1, None, None] + t * dw[..., 1]
                jx = ((torch.remainder(k * 0.6180339887, 1.0) - 0.5) * 0.6)[:, None, None]    # per-board lateral jitter
                inside = (hx - jx > xa) & (hx - jx < xb) & (hy > ya) & (hy < yb)
                kk = k[:, None, None]
                take(inside & (t > 0.05) & (t < best_t), t, hx * texels_per_m + 211.0 * bi + 97.0 * kk,
                     hy * texels_per_m + 59.0 * bi + 31.0 * kk, (13.0 * bi + 7.0 * kk).expand_as(t), torch.full_like(t, 5.0 * bi))
        # pass 2: one texture lookup per pixel -- base layer + near-range detail layer, faded with distance
        wd = (_DETAIL_GAIN * (1.0 - best_t / detail_range)).clamp(min=0.0)
        val = sample(b_tu, b_tv) + wd * (sample(b_tu * _DETAIL_FREQ + b_du, b_tv * _DETAIL_FREQ + b_dv) - 128.0)
        fade = torch.exp(-torch.clamp(best_t - 18.0, min=0.0) / 10.0)
        val = torch.where(best_t < 1e8, 128.0 + (val - 128.0) * fade, torch.full_like(val, 128.0))
        out[s:s + chunk] = val.round().clamp(0, 255).to(torch.uint8)
    return out


def make_sequence(width, height, nframes, device="cpu", scene_seed=1234, traj_seed=5678):
    """Convenience: (frames u8 tensor (n,H,W), R_wc, C_w, K)."""
    tex = make_texture(seed=scene_seed)
    R_wc, C_w = make_trajectory(nframes, seed=traj_seed)
    frames = render_frames(width, height, R_wc, C_w, tex, device=device)
    return frames, R_wc, C_w, camera_matrix(width, height)
