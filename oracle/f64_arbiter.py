"""Float64 arbiter of the rasterizer's gradients -- TEST INFRASTRUCTURE ONLY (never imported by the product).

SURVEY.md §7 "hard parts": the reference accumulates its gradients with float atomics in unspecified order, so two runs of
the REFERENCE differ by up to ~7e-4 (norm-wise) in dL_drotations at 1 M Gaussians.  Comparing two float32 implementations
with each other therefore cannot resolve north_star's 1e-3 gate; a double-precision evaluation of the same function can:
the gate becomes  |ours - f64| <= max(1e-3 |f64|, 1.5 |reference - f64|)  per gradient tensor.

This module re-states the differentiable part of the forward pass in torch float64 and lets autograd produce the
gradients (an independent derivation: nothing of the hand-written backward kernels or of oracle/dqo_oracle.c's backward is
reused).  Formulas follow the reference sources:
  projection, EWA covariance, conic        RAST/cuda_rasterizer/forward.cu:158-197, 238-354, auxiliary.h:39-97
  3D covariance from scale / quaternion    forward.cu:202-235  (quaternion NOT re-normalised)
  SH -> RGB with clamping                  forward.cu:104-155
  alpha blending                           forward.cu:757-848  (power > 0 and alpha < 1/255 rejections, min(0.99, .))
  hit depth: plane intersection / centre   forward.cu:779-810  (argmin-scale axis of R as the surfel normal)
  clamped-frustum gradient convention      backward.cu:300-312 (x_grad_mul / y_grad_mul: a clamped t.x is a constant)
The DISCRETE decisions (depth-sorted per-tile lists, last contributor per pixel, hit Gaussian per pixel, radii > 0) are
inputs: they come from a float32 forward (this library's, bit-identical to the reference's in every integer artefact),
so that the arbiter differentiates the same piecewise-smooth branch as the implementations under test.
"""
import math

import torch

SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
SH_C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
         1.445305721320277, -0.5900435899266435]


def _rotation(q):
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    return torch.stack([
        torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)], -1),
        torch.stack([2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)], -1),
        torch.stack([2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], -1)], 1)  # [n,3,3] row-major


def _sh_rgb(deg, sh, dirs):
    x, y, z = dirs[:, 0:1], dirs[:, 1:2], dirs[:, 2:3]
    res = SH_C0 * sh[:, 0]
    if deg > 0:
        res = res - SH_C1 * y * sh[:, 1] + SH_C1 * z * sh[:, 2] - SH_C1 * x * sh[:, 3]
        if deg > 1:
            xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
            res = (res + SH_C2[0] * xy * sh[:, 4] + SH_C2[1] * yz * sh[:, 5] + SH_C2[2] * (2 * zz - xx - yy) * sh[:, 6]
                   + SH_C2[3] * xz * sh[:, 7] + SH_C2[4] * (xx - yy) * sh[:, 8])
            if deg > 2:
                res = (res + SH_C3[0] * y * (3 * xx - yy) * sh[:, 9] + SH_C3[1] * xy * z * sh[:, 10]
                       + SH_C3[2] * y * (4 * zz - xx - yy) * sh[:, 11] + SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[:, 12]
                       + SH_C3[4] * x * (4 * zz - xx - yy) * sh[:, 13] + SH_C3[5] * z * (xx - yy) * sh[:, 14]
                       + SH_C3[6] * x * (xx - 3 * yy) * sh[:, 15])
    return torch.clamp_min(res + 0.5, 0.0)  # clamped channels get zero gradient (forward.cu:151-153, backward.cu:166-170)


def gradients(scene, lists, grad_color, grad_depth, device="cpu", tile_chunk=8):
    """scene: dict of float32 tensors xyz [P,3], scales [P,3], rotations [P,4], opacity [P,1], shs [P,M,3] or rgb [P,3]
    (colors_precomp), view [4,4], proj [4,4] (both column-major as the kernels take them = transposed W2C / full proj),
    campos [3], bg [3], and scalars W, H, tanfovx, tanfovy, cx, cy, sh_degree, scale_modifier, opaque_threshold,
    depth_threshold, normal_threshold.
    lists: dict of integer artefacts of a float32 forward: point_list [R] (Gaussian ids, tile-major, depth order),
    ranges [T,2], n_contrib [H,W], hit [H,W] (depth index map, -1 none), radii [P].
    Returns float64 gradients of sum(grad_color * color) + sum(grad_depth * depth) w.r.t. the inputs, as a dict with the
    reference's names (dL_dmeans3D, dL_dsh / dL_dcolors, dL_dopacity, dL_dscales, dL_drotations)."""
    dd = dict(dtype=torch.float64, device=device)
    W, H = int(scene["W"]), int(scene["H"])
    fx, fy = W / (2.0 * scene["tanfovx"]), H / (2.0 * scene["tanfovy"])
    cx, cy = float(scene["cx"]), float(scene["cy"])
    leaf = {k: scene[k].to(**dd).clone().requires_grad_(True) for k in ("xyz", "scales", "rotations", "opacity")}
    use_sh = scene.get("shs") is not None
    leaf["col"] = (scene["shs"] if use_sh else scene["rgb"]).to(**dd).clone().requires_grad_(True)
    view = scene["view"].to(**dd).reshape(4, 4)   # column-major storage: view[c, r] = W2C[r, c]
    proj = scene["proj"].to(**dd).reshape(4, 4)
    W2C, PROJ = view.t(), proj.t()
    campos, bg = scene["campos"].to(**dd), scene["bg"].to(**dd)
    vis = (lists["radii"].to(device) > 0).nonzero().squeeze(1)
    slot = torch.full((scene["xyz"].shape[0],), -1, dtype=torch.long, device=device)
    slot[vis] = torch.arange(vis.numel(), device=device)

    # ---- per-Gaussian quantities (graph kept) ----
    p = leaf["xyz"][vis]
    ph = torch.cat([p, torch.ones_like(p[:, :1])], 1)
    hom = ph @ PROJ.t()
    p_w = 1.0 / (hom[:, 3] + 0.0000001)
    ppx, ppy = hom[:, 0] * p_w, hom[:, 1] * p_w
    pc = ph @ W2C.t()                       # camera-space centre
    R = _rotation(leaf["rotations"][vis])
    s = leaf["scales"][vis] * float(scene["scale_modifier"])
    Sigma = (R * (s * s)[:, None, :]) @ R.transpose(1, 2)
    tz = pc[:, 2]
    limx, limy = 1.3 * scene["tanfovx"], 1.3 * scene["tanfovy"]
    txtz, tytz = pc[:, 0] / tz, pc[:, 1] / tz
    tx = torch.where((txtz < -limx) | (txtz > limx), (txtz.clamp(-limx, limx) * tz).detach(), pc[:, 0])
    ty = torch.where((tytz < -limy) | (tytz > limy), (tytz.clamp(-limy, limy) * tz).detach(), pc[:, 1])
    zero = torch.zeros_like(tz)
    J = torch.stack([torch.stack([fx / tz, zero, -fx * tx / (tz * tz)], -1),
                     torch.stack([zero, fy / tz, -fy * ty / (tz * tz)], -1)], 1)  # [n,2,3]
    Tm = J @ W2C[:3, :3]
    cov = Tm @ Sigma @ Tm.transpose(1, 2)
    a_, b_, c_ = cov[:, 0, 0] + 0.3, cov[:, 0, 1], cov[:, 1, 1] + 0.3
    det = a_ * c_ - b_ * b_
    conic = torch.stack([c_ / det, -b_ / det, a_ / det], -1)
    mean2d = torch.stack([ppx * W * 0.5 + cx, ppy * H * 0.5 + cy], -1)
    if use_sh:
        d = p - campos
        rgb = _sh_rgb(int(scene["sh_degree"]), leaf["col"][vis], d / d.norm(dim=1, keepdim=True))
    else:
        rgb = leaf["col"][vis]
    opac = leaf["opacity"][vis, 0]
    # surfel normal = column (argmin scale) of R, in camera space; reference ties: first minimum (forward.cu:20-35)
    raw_s = leaf["scales"][vis]
    ax = torch.argmin(raw_s.detach(), dim=1)
    n_w = torch.gather(R, 2, ax[:, None, None].expand(-1, 3, 1)).squeeze(2)
    n_c = n_w @ W2C[:3, :3].t()
    smax = raw_s.detach().max(dim=1).values * float(scene["scale_modifier"])
    Q = torch.cat([mean2d, conic, opac[:, None], rgb, pc[:, :3], n_c], 1)  # [n, 2+3+1+3+3+3 = 15]
    Qacc = torch.zeros_like(Q)

    # ---- per-tile blending on detached copies, gradients accumulated into Qacc ----
    ranges = lists["ranges"].to(device).long()
    plist = lists["point_list"].to(device).long()
    ncon = lists["n_contrib"].to(device).long()
    hit = lists["hit"].to(device).long()
    gcol, gdep = grad_color.to(**dd), grad_depth.to(**dd).reshape(H, W)
    tiles_x = (W + 15) // 16
    T_tiles = ranges.shape[0]
    for t0 in range(T_tiles):
        lo, hi = int(ranges[t0, 0]), int(ranges[t0, 1])
        if hi <= lo:
            continue
        ty0, tx0 = (t0 // tiles_x) * 16, (t0 % tiles_x) * 16
        ys = torch.arange(ty0, min(ty0 + 16, H), device=device)
        xs = torch.arange(tx0, min(tx0 + 16, W), device=device)
        py, px = torch.meshgrid(ys, xs, indexing="ij")
        py, px = py.reshape(-1), px.reshape(-1)
        nc = ncon[py, px]
        L = int(nc.max())
        if L == 0 and int((hit[py, px] >= 0).sum()) == 0:
            continue
        ids = slot[plist[lo:lo + max(L, 1)]]
        q = Q[ids].detach().requires_grad_(True)          # [L, 15]
        dx = q[None, :, 0] - px[:, None].to(**dd)
        dy = q[None, :, 1] - py[:, None].to(**dd)
        power = -0.5 * (q[None, :, 2] * dx * dx + q[None, :, 4] * dy * dy) - q[None, :, 3] * dx * dy
        alpha = torch.clamp_max(q[None, :, 5] * torch.exp(power), 0.99)
        pos = torch.arange(ids.numel(), device=device)[None, :]
        valid = (power <= 0) & (alpha >= 1.0 / 255.0) & (pos < nc[:, None])
        a_eff = torch.where(valid, alpha, torch.zeros_like(alpha))
        one_m = 1.0 - a_eff
        T_excl = torch.cumprod(torch.cat([torch.ones_like(one_m[:, :1]), one_m[:, :-1]], 1), 1)
        w = a_eff * T_excl
        color = w @ q[:, 6:9] + (T_excl[:, -1] * one_m[:, -1])[:, None] * bg[None, :]
        loss = (color * gcol[:, py, px].t()).sum()
        # depth of the hit Gaussian (forward.cu:779-810); the hit id is part of `lists`
        hid = hit[py, px]
        hp = (hid >= 0).nonzero().squeeze(1)
        if hp.numel():
            hs = slot[hid[hp]]
            # rows of Q for the hit Gaussians, attached to the same leaf through an index into q when they are in the
            # staged prefix (they always are: the hit happens at or before the last contributor... except when the pixel
            # terminates on the hit itself) -- use a separate detached leaf to stay general
            qh = Q[hs].detach().requires_grad_(True)
            pcx, ncx = qh[:, 9:12], qh[:, 12:15]
            rx = (px[hp].to(**dd) - cx) / fx
            ry = (py[hp].to(**dd) - cy) / fy
            inv = 1.0 / torch.sqrt(rx * rx + ry * ry + 1.0)
            ray = torch.stack([rx * inv, ry * inv, inv], -1)
            den = (ray * ncx).sum(-1)
            tt = (pcx * ncx).sum(-1) / (den + 1e-8)
            hz = tt * ray[:, 2]
            plane = ((hz - pcx[:, 2]).abs() <= smax[hs] * float(scene["depth_threshold"])) & (den.abs() >= float(scene["normal_threshold"]))
            depth = torch.where(plane, hz, pcx[:, 2])
            loss = loss + (depth * gdep[py[hp], px[hp]]).sum()
            loss.backward()
            Qacc.index_add_(0, hs, qh.grad)
        else:
            loss.backward()
        Qacc.index_add_(0, ids, q.grad)
    Q.backward(Qacc)
    out = {"dL_dmeans3D": leaf["xyz"].grad, "dL_dopacity": leaf["opacity"].grad, "dL_dscales": leaf["scales"].grad,
           "dL_drotations": leaf["rotations"].grad}
    out["dL_dsh" if use_sh else "dL_dcolors"] = leaf["col"].grad
    return {k: (v if v is not None else torch.zeros(1, **dd)) for k, v in out.items()}
