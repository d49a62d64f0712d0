"""Mapping-step operators: masked L1 colour/depth loss, fused multi-tensor Adam, and the iteration that
`Mapping.local_optimize` / `loss_update` run (reference SLAM/multiprocess/mapper.py:531-605, 799-928).

The reference has no native boundary here (inline PyTorch); these are opt-in replacements with the same
semantics: `FusedAdam` accepts the param-group dicts of `GaussianPointCloud.parametrize`
(SLAM/gaussian_pointcloud.py:331-378) and behaves like `torch.optim.Adam(l, lr=0.0, eps=1e-15)`.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import AdamTensor, check, lib, ptr


def _stream():
    return torch.cuda.current_stream().cuda_stream


class _MaskedL1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, depth, hit_depth, gt_color, gt_depth, render_mask, color_weight, depth_weight,
                depth_err_thres):
        L = lib()
        dev = image.device
        H, W = image.shape[1], image.shape[2]
        image_c, depth_c = image.contiguous(), depth.contiguous()
        mask_c = None
        if render_mask is not None:
            mask_c = render_mask.contiguous()
            if mask_c.dtype == torch.bool:
                mask_c = mask_c.view(torch.uint8)
            elif mask_c.dtype != torch.uint8:
                mask_c = (mask_c != 0).view(torch.uint8)
        d_img = torch.empty_like(image_c)
        d_depth = torch.empty_like(depth_c)
        out = torch.empty((4,), dtype=torch.float32, device=dev)
        counts = torch.empty((2,), dtype=torch.int32, device=dev)
        ws = torch.empty((L.dqo_loss_workspace_bytes(W, H),), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            check(L.dqo_masked_l1_loss(W, H, ptr(image_c), ptr(depth_c), ptr(hit_depth.contiguous()),
                                       ptr(gt_color.contiguous()), ptr(gt_depth.contiguous()), ptr(mask_c),
                                       float(color_weight), float(depth_weight), float(depth_err_thres), ptr(d_img),
                                       ptr(d_depth), ptr(out), ptr(counts), ptr(ws), _stream()), "dqo_masked_l1_loss")
        ctx.save_for_backward(d_img, d_depth)
        ctx.mark_non_differentiable(counts)
        return out[0], out[1], out[2], counts

    @staticmethod
    def backward(ctx, g_total, g_color, g_depth, _):
        d_img, d_depth = ctx.saved_tensors
        # the component losses are reported only (detached), like the `.item()` reports of the reference
        return d_img * g_total, d_depth * g_total, None, None, None, None, None, None, None


def masked_l1_loss(image, depth, hit_depth, gt_color, gt_depth, render_mask=None, color_weight=0.8, depth_weight=1.0,
                   depth_err_thres=0.1):
    """total = depth_weight * mean|d - gt_d|[valid] + color_weight * mean|img - gt|[mask]  (mapper.py:847-875).

    image [3,H,W], depth [1,H,W], hit_depth [1,H,W] int32 (the rasterizer's depth index map), gt_color [H,W,3],
    gt_depth [H,W,1] or [H,W], render_mask [H,W] bool or None.  Returns (total, colour, depth, counts[2])."""
    return _MaskedL1.apply(image, depth, hit_depth, gt_color, gt_depth, render_mask, color_weight, depth_weight,
                           depth_err_thres)


class _SSIMLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, gt_color):
        L = lib()
        dev = image.device
        H, W = image.shape[1], image.shape[2]
        image_c = image.contiguous()
        d_img = torch.empty_like(image_c)
        out = torch.empty((2,), dtype=torch.float32, device=dev)
        ws = torch.empty((L.dqo_ssim_workspace_bytes(W, H),), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            check(L.dqo_ssim_loss(W, H, ptr(image_c), ptr(gt_color.contiguous()), 1.0, ptr(d_img), 0, ptr(out), ptr(ws),
                                  _stream()), "dqo_ssim_loss")
        ctx.save_for_backward(d_img)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        (d_img,) = ctx.saved_tensors
        return d_img * g, None


def ssim_loss(image, gt_color):
    """1 - ssim(image, gt) as `Mapping.loss_update` adds it in the mask-less global pass (mapper.py:839-841;
    utils/loss_utils.py:61-99: 11x11 Gaussian window, zero padding, mean over 3*H*W), value and image gradient in two
    launches.  image [3,H,W] (differentiable), gt_color [H,W,3] float32 CUDA tensors.  Returns a device scalar."""
    if image.dim() != 3 or image.shape[0] != 3 or tuple(gt_color.shape) != (image.shape[1], image.shape[2], 3):
        raise ValueError("ssim_loss expects image [3,H,W] and gt_color [H,W,3]")
    return _SSIMLoss.apply(image, gt_color)


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam semantics (amsgrad=False, weight_decay=0, maximize=False) in ONE launch over all groups."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, confidence=None, confidence_param=None):
        defaults = dict(lr=lr, betas=betas, eps=eps)
        super().__init__(params, defaults)
        self.confidence = confidence          # [P] or [P,1] float tensor bumped where grad(confidence_param) != 0
        self.confidence_param = confidence_param
        self._step = 0

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        arr = (AdamTensor * _lib.ADAM_MAX_TENSORS)()
        n, conf_idx, keep = 0, -1, []
        betas, eps = None, None
        for group in self.param_groups:
            if betas is None:
                betas, eps = group["betas"], group["eps"]
            elif betas != group["betas"] or eps != group["eps"]:
                raise ValueError("FusedAdam requires identical betas/eps across param groups")
            for p in group["params"]:
                if p.grad is None:
                    continue
                if n >= _lib.ADAM_MAX_TENSORS:
                    raise ValueError("FusedAdam supports at most %d tensors" % _lib.ADAM_MAX_TENSORS)
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise ValueError("FusedAdam expects contiguous float32 CUDA parameters")
                st = self.state[p]
                if len(st) == 0:
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                g = p.grad.contiguous()
                keep.append(g)
                width = int(p.numel() // p.shape[0]) if p.dim() > 0 and p.shape[0] > 0 else 0
                arr[n] = AdamTensor(ptr(p), ptr(g), ptr(st["exp_avg"]), ptr(st["exp_avg_sq"]), p.numel(),
                                    float(group["lr"]), width, 0)
                if self.confidence_param is not None and p is self.confidence_param:
                    conf_idx = n
                n += 1
        if n == 0:
            return loss
        self._step += 1
        dev = keep[0].device
        conf = self.confidence if conf_idx >= 0 else None
        with torch.cuda.device(dev):
            check(lib().dqo_adam_step(arr, n, self._step, float(betas[0]), float(betas[1]), float(eps), ptr(conf),
                                      conf_idx, _stream()), "dqo_adam_step")
        return loss


class MappingStep:
    """One iteration of the reference's mapping hot loop (mapper.py:568-599 + loss_update :799-928) with the
    B200 operators plugged in: activations (torch) -> rasterize (C-ABI) -> masked L1 loss (C-ABI) -> backward
    (C-ABI + torch activations) -> fused Adam (C-ABI) with the confidence bump.  Optional terms as in loss_update: with
    `ssim_weight` > 0 calls without a render mask add ssim_weight * (1 - ssim) (mapper.py:839-841, `ssim_loss`); when
    `params` holds "semantics" [P,3] and the call passes `gt_semantic`, semantic_weight * masked L1 of the semantic image
    (mapper.py:877-880), rendered and back-propagated over the lists of the main render (`blend_extra_colors_grad`)."""

    def __init__(self, params, lrs, settings_fn, color_weight=0.8, depth_weight=1.0, depth_err_thres=0.1,
                 confidence=None, optimizer="fused", ssim_weight=0.0, semantic_weight=0.1):
        # params: dict with raw leaf tensors xyz[P,3], f_dc[P,1,3], f_rest[P,M-1,3], opacity[P,1], scaling[P,3], rotation[P,4]
        # and optionally semantics[P,3] (parameter group "semantics_color", gaussian_pointcloud.py:371-378)
        self.p = params
        self.ssim_w, self.sem_w = float(ssim_weight), float(semantic_weight)
        groups = [{"params": [params[k]], "lr": lrs[k], "name": k}
                  for k in ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")]
        if params.get("semantics") is not None:
            groups.append({"params": [params["semantics"]], "lr": lrs["semantics"], "name": "semantics_color"})
        if optimizer == "fused":
            self.opt = FusedAdam(groups, lr=0.0, eps=1e-15, confidence=confidence, confidence_param=params["f_dc"])
        else:
            self.opt = torch.optim.Adam(groups, lr=0.0, eps=1e-15)
        self.confidence = confidence
        self.fused = optimizer == "fused"
        self.settings_fn = settings_fn
        self.cw, self.dw, self.thr = color_weight, depth_weight, depth_err_thres

    def activated(self):
        p = self.p
        return dict(xyz=p["xyz"], opacity=torch.sigmoid(p["opacity"]), scales=torch.exp(p["scaling"]),
                    rotations=torch.nn.functional.normalize(p["rotation"]),
                    shs=torch.cat((p["f_dc"], p["f_rest"]), dim=1))

    def __call__(self, frame, tile_mask, gt_color, gt_depth, render_mask, gt_semantic=None):
        from .rasterizer import GaussianRasterizer, _RasterizeGaussians, blend_extra_colors_grad
        a = self.activated()
        rs = self.settings_fn(frame)
        rast = GaussianRasterizer(rs)
        color, depth, hit_color, hit_depth, hcw, hdw, T_map, n_touched, radii = rast(
            means3D=a["xyz"], opacities=a["opacity"], shs=a["shs"], scales=a["scales"], rotations=a["rotations"],
            tile_mask=tile_mask)
        total, lc, ld, counts = masked_l1_loss(color, depth, hit_depth, gt_color, gt_depth, render_mask, self.cw,
                                               self.dw, self.thr)
        self.ssim_value = self.semantic_value = None
        if render_mask is None and self.ssim_w > 0:
            self.ssim_value = ssim_loss(color, gt_color)
            total = total + self.ssim_w * self.ssim_value
        if gt_semantic is not None and self.p.get("semantics") is not None:
            sem = blend_extra_colors_grad(_RasterizeGaussians.last_state, rs, self.p["semantics"], a["xyz"], a["opacity"],
                                          a["scales"], a["rotations"]).permute(1, 2, 0)
            m = render_mask.bool() if render_mask is not None else torch.ones(sem.shape[:2], dtype=torch.bool, device=sem.device)
            self.semantic_value = torch.abs(sem[m] - gt_semantic[m]).mean()
            total = total + self.sem_w * self.semantic_value
        total.backward()
        self.opt.step()
        if not self.fused and self.confidence is not None:
            grad_mask = (self.p["f_dc"].grad.abs() != 0).any(dim=-1)
            self.confidence[grad_mask.view(-1)] += 1
        self.opt.zero_grad(set_to_none=True)
        return total.detach(), lc.detach(), ld.detach()


class FusedMappingStep:
    """The iteration of `Mapping.local_optimize` (mapper.py:568-599 + loss_update :799-928) enqueued by ONE C-ABI call
    (`dqo_mapping_step`) with no host synchronisation: activations, rasterize forward, masked L1 colour + depth loss,
    backward, activation backward, the attach term and Adam all run inside the library on the raw parameter tensors,
    which are updated in place (opt-in).

    Loss terms: masked L1 colour, masked L1 depth, after `begin_window(attach=True)` the attach term (mapper.py:810-829),
    with `ssim_weight` > 0 the SSIM term -- like the reference only for calls without a render_mask (:839-841, the final
    global pass) -- and, when `params` holds "semantics" and the call passes `gt_semantic`, the semantic colour term
    `semantic_weight` * masked L1 of the semantic image (:877-880), whose gradient updates the semantic colours (group
    "semantics_color", lrs["semantics"]) and the geometry.  NOT covered: the normal term (weight 0 in every shipped
    config).  The instance term (:881-902, Method 2) is an L1 on T_map, whose gradient the rasterizer's backward drops
    (diff_gaussian_rasterization_depth/__init__.py:184): it never changes a parameter and is not evaluated here.
    `loss` holds {total, colour, depth, attach, 1 - ssim, semantic L1, 0, 0}.

    params: dict of raw contiguous float32 CUDA tensors xyz [P,3], f_dc [P,1,3], f_rest [P,M-1,3], opacity [P,1],
    scaling [P,3], rotation [P,4] and optionally semantics [P,3] (the tensors of GaussianPointCloud.parametrize); lrs:
    dict name -> lr (re-read on every call).  `__call__` returns device tensors (total, colour, depth loss); reading them is the only
    synchronisation.  The Adam step number lives on the device; `check()` reports how many steps the device skipped
    since the last check (instance overflow) -- the counter is sticky, so checking once after a window is enough.
    `graph(...)` captures the step of one keyframe into a CUDA graph (one launch per iteration instead of ~45)."""

    ORDER = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")

    def __init__(self, params, lrs, width, height, color_weight=0.8, depth_weight=1.0, depth_err_thres=0.1,
                 confidence=None, betas=(0.9, 0.999), eps=1e-15, capacity=None, need_n_touched=True,
                 front_instances=0, back_instances=0, ssim_weight=0.0, semantic_weight=0.1):
        L = lib()
        self.p = params
        self.ssim_w, self.sem_w = float(ssim_weight), float(semantic_weight)
        self.has_sem = params.get("semantics") is not None and params["semantics"].numel() > 1
        # the step workspace carries the scratch of the optional terms only when they can occur
        self.terms = (_lib.STEP_TERM_SSIM if self.ssim_w > 0 else 0) | (_lib.STEP_TERM_SEMANTIC if self.has_sem else 0)
        self.front, self.back = int(front_instances), int(back_instances)
        for k in self.ORDER:
            t = params[k]
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise ValueError("FusedMappingStep expects contiguous float32 CUDA tensors (%s)" % k)
        self.dev = params["xyz"].device
        self.P = params["xyz"].shape[0]
        self.M = 1 + (params["f_rest"].shape[1] if params["f_rest"].numel() else 0)
        if self.M not in (1, 16):
            raise ValueError("FusedMappingStep supports SH storage of 1 or 16 coefficients per channel")
        self.W, self.H = int(width), int(height)
        self.lrs = lrs
        self.betas, self.eps = betas, eps
        self.cw, self.dw, self.thr = float(color_weight), float(depth_weight), float(depth_err_thres)
        self.confidence = confidence
        self.need_n_touched = need_n_touched
        self.state = {k: (torch.zeros_like(params[k]), torch.zeros_like(params[k])) for k in self.ORDER}
        if self.has_sem:
            t = params["semantics"]
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == (self.P, 3)):
                raise ValueError("FusedMappingStep expects semantics as a contiguous float32 CUDA tensor [P,3]")
            self.state["semantics"] = (torch.zeros_like(t), torch.zeros_like(t))
        # one byte per Gaussian: has it ever received a non-zero gradient?  (zero-initialised with the moments; Gaussians
        # that never have are skipped by the backward's gradient writes and by the optimiser, see dqo_map_params.ever)
        self.ever = torch.zeros(self.P, dtype=torch.uint8, device=self.dev)
        # compact list of the flagged Gaussians + its length: the optimiser walks the list, not the cloud
        self.ever_list = torch.zeros(self.P, dtype=torch.int32, device=self.dev)
        self.ever_count = torch.zeros(1, dtype=torch.int32, device=self.dev)
        # [0] Adam steps taken, [1] steps skipped by the device (overflow), sticky until check()
        self.step_state = torch.zeros(4, dtype=torch.int32, device=self.dev)
        self.init = None            # attach reference (history_stat of local_optimize), see begin_window
        self.attach_count = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.attach_weight, self.attach_thres = 1000.0, 0.9
        self._kf_cache = {}
        self.capacity = int(capacity) if capacity else max(8 * self.P, 1 << 16)
        self._alloc()
        self.loss = torch.zeros(8, dtype=torch.float32, device=self.dev)
        self._loss_views = (self.loss[0], self.loss[1], self.loss[2])
        self.counts = torch.zeros(2, dtype=torch.int32, device=self.dev)
        self.status = torch.zeros(_lib.ST_WORDS, dtype=torch.int32, device=self.dev)

    @property
    def step(self):
        """Adam steps taken so far (device counter; one host synchronisation)."""
        return int(self.step_state[0].item())

    def begin_window(self, attach=True, attach_weight=1000.0, opacity_thres=0.9):
        """Start of an optimisation window, as `Mapping.local_optimize` does it (mapper.py:531-548): a fresh Adam (zero
        moments, step 0) and, with attach=True, `history_stat` = the current raw parameters as the anchor of the attach
        term for Gaussians whose initial opacity is below `opacity_thres` (mapper.py:810-829)."""
        for m, v in self.state.values():
            m.zero_()
            v.zero_()
        self.ever.zero_()
        self.ever_count.zero_()
        self.step_state.zero_()
        self._mp_key = None
        if not attach:
            self.init = None
            return
        self.init = {k: self.p[k].detach().clone() for k in ("xyz", "scaling", "rotation", "opacity")}
        self.attach_weight, self.attach_thres = float(attach_weight), float(opacity_thres)
        with torch.cuda.device(self.dev):
            check(lib().dqo_attach_count(self.P, ptr(self.init["opacity"]), self.attach_thres, ptr(self.attach_count),
                                         _stream()), "dqo_attach_count")

    def attach_loss(self):
        """Value of the attach term at the parameters the last step started from (device scalar; the reference reports it
        as `scale_loss`, mapper.py:919)."""
        return self.loss[3]

    def _map_params(self):
        """ctypes view of the parameter / moment tensors; rebuilt when a tensor was replaced or a learning rate changed."""
        lrs = tuple(float(self.lrs[k]) for k in self.ORDER)
        if self.has_sem:
            lrs = lrs + (float(self.lrs["semantics"]), self.p["semantics"].data_ptr())
        key = tuple(self.p[k].data_ptr() for k in self.ORDER) + lrs + (
            self.confidence.data_ptr() if self.confidence is not None else 0,
            self.init["xyz"].data_ptr() if self.init is not None else 0, self.attach_weight, self.attach_thres)
        if getattr(self, "_mp_key", None) != key:
            for k in self.ORDER:
                if not self.p[k].is_contiguous():
                    raise ValueError("FusedMappingStep expects contiguous parameters (%s)" % k)
            mp = _lib.MapParams()
            for i, k in enumerate(self.ORDER):
                mp.param[i] = ptr(self.p[k])
                mp.exp_avg[i] = ptr(self.state[k][0])
                mp.exp_avg_sq[i] = ptr(self.state[k][1])
                mp.lr[i] = lrs[i]
            mp.confidence = ptr(self.confidence)
            mp.ever = ptr(self.ever)
            mp.ever_list, mp.ever_count = ptr(self.ever_list), ptr(self.ever_count)
            if self.init is not None:
                mp.init_xyz, mp.init_scaling = ptr(self.init["xyz"]), ptr(self.init["scaling"])
                mp.init_rotation, mp.init_opacity = ptr(self.init["rotation"]), ptr(self.init["opacity"])
                mp.attach_count = ptr(self.attach_count)
            mp.attach_weight, mp.attach_opacity_thres = self.attach_weight, self.attach_thres
            mp.step_state = ptr(self.step_state)
            mp.workspace_terms = self.terms
            if self.has_sem:
                mp.semantics = ptr(self.p["semantics"])
                mp.semantics_exp_avg, mp.semantics_exp_avg_sq = (ptr(t) for t in self.state["semantics"])
                mp.lr_semantics = float(self.lrs["semantics"])
            self._mp, self._mp_key = mp, key
        return self._mp

    def _alloc(self):
        n = lib().dqo_mapping_step_workspace_bytes(self.P, self.M, self.W, self.H, self.capacity, self.terms)
        if n == 0:
            raise _lib.DqoError("workspace size query failed: %s" % lib().dqo_last_error().decode())
        self.ws = torch.empty((n,), dtype=torch.uint8, device=self.dev)
        # persistent workspace: gradient accumulators cleared once, kept clean by every step (settings.geom_clean)
        with torch.cuda.device(self.dev):
            check(lib().dqo_mapping_step_workspace_init(self.P, self.M, self.W, self.H, self.capacity, self.terms, ptr(self.ws),
                                                        _stream()),
                  "dqo_mapping_step_workspace_init")

    def _keyframe(self, rs, tile_mask, gt_color, gt_depth, render_mask, gt_semantic=None):
        from .rasterizer import _make_settings
        # the ctypes views of a keyframe (settings + pointers) are cached per keyframe: a mapping window revisits the
        # same few keyframes, and building the structs costs more host time than launching the step.  Inputs must be
        # contiguous (no hidden copies that an in-place update of the source would miss).
        key = (id(rs), tile_mask.data_ptr(), gt_color.data_ptr(), gt_depth.data_ptr(),
               render_mask.data_ptr() if render_mask is not None else 0, self.front, self.back, self.need_n_touched,
               self.cw, self.dw, self.thr, self.ssim_w, self.sem_w, gt_semantic.data_ptr() if gt_semantic is not None else 0)
        hit = self._kf_cache.get(key)
        if hit is None:
            mask = render_mask
            if mask is not None and mask.dtype == torch.bool:
                mask = mask.view(torch.uint8)
            if gt_semantic is not None and not self.has_sem:
                raise ValueError("gt_semantic given but params has no 'semantics' tensor")
            for name, t in (("gt_color", gt_color), ("gt_depth", gt_depth), ("render_mask", mask), ("tile_mask", tile_mask),
                            ("gt_semantic", gt_semantic)):
                if t is not None and not t.is_contiguous():
                    raise ValueError("FusedMappingStep expects a contiguous %s" % name)
            if mask is not None and mask.dtype != torch.uint8:
                raise TypeError("render_mask must be a bool or uint8 tensor")
            tensors = (gt_color, gt_depth, mask, tile_mask, rs, gt_semantic)  # kept alive with the cache entry
            kf = _lib.Keyframe(ptr(gt_color), ptr(gt_depth), ptr(mask), ptr(tile_mask), ptr(rs.viewmatrix),
                               ptr(rs.projmatrix), ptr(rs.campos), ptr(rs.bg), self.cw, self.dw, self.thr, self.ssim_w,
                               self.sem_w if gt_semantic is not None else 0.0, ptr(gt_semantic))
            s = _make_settings(self.P, int(rs.sh_degree), self.M, self.W, self.H, rs.tanfovx, rs.tanfovy, rs.cx, rs.cy,
                               rs.scale_modifier, rs.color_sigma, rs.opaque_threshold, rs.depth_threshold,
                               rs.normal_threshold, rs.T_threshold, rs.prefiltered, rs.debug, self.need_n_touched,
                               self.front, self.back, geom_clean=True)
            if len(self._kf_cache) >= 256:
                self._kf_cache.clear()
            hit = self._kf_cache[key] = (s, kf, tensors)
        return hit[0], hit[1]

    def __call__(self, rs, tile_mask, gt_color, gt_depth, render_mask=None, gt_semantic=None):
        """rs: GaussianRasterizationSettings of the keyframe (as built by SLAM/render.py:142-162); gt_semantic: [H,W,3]
        image_input["semantics_color"] (enables the semantic term) or None."""
        mp = self._map_params()
        s, kf = self._keyframe(rs, tile_mask, gt_color, gt_depth, render_mask, gt_semantic)
        args = (s, mp, kf, 0, float(self.betas[0]), float(self.betas[1]), float(self.eps), ptr(self.ws),
                self.capacity, ptr(self.loss), ptr(self.counts), ptr(self.status))
        if torch.cuda.current_device() == self.dev.index:  # the common case: no device switch on the hot path
            check(lib().dqo_mapping_step(*args, _stream()), "dqo_mapping_step")
        else:
            with torch.cuda.device(self.dev):
                check(lib().dqo_mapping_step(*args, _stream()), "dqo_mapping_step")
        return self._loss_views

    def optimize_window(self, keyframes, iters, rng=None, attach=True):
        """The iteration loop of `Mapping.local_optimize` (mapper.py:531-599) on this step: `begin_window` (fresh Adam,
        attach anchors), then `iters` iterations, each on a randomly chosen keyframe of the window during the first half
        and on the newest keyframe afterwards (:573-575).  keyframes: sequence of dicts with rs, tile_mask, gt_color,
        gt_depth and optionally render_mask / gt_semantic (the tensors of `processed_frames` / `processed_map` and the masks
        of `evaluate_render_range`).  Nothing synchronises inside the loop; the overflow counter is read once at its end:
        steps the device skipped because a keyframe outgrew the instance buffers are repeated (on the newest keyframe, like
        the second half of the window) after the workspace has been re-sized for what the device reported.  Returns the
        keyframe index of every iteration that was enqueued."""
        import random
        rng = rng or random
        self.begin_window(attach=attach)
        used = []

        def run(n, first):
            for it in range(first, first + n):
                idx = rng.randint(0, len(keyframes) - 1)
                if it > iters / 2:
                    idx = len(keyframes) - 1
                kf = keyframes[idx]
                self(kf["rs"], kf["tile_mask"], kf["gt_color"], kf["gt_depth"], kf.get("render_mask"), kf.get("gt_semantic"))
                used.append(idx)

        run(iters, 0)
        for _ in range(6):
            skipped = self.check(auto_resize=True)
            if not isinstance(skipped, int):    # the status words: nothing was skipped
                return used
            run(skipped, iters)
        raise _lib.DqoError("optimize_window: instance buffers still overflow after six re-sizes")

    def graph(self, rs, tile_mask, gt_color, gt_depth, render_mask=None, warmup=True, gt_semantic=None):
        """CUDA graph of this keyframe's step: `g = step.graph(...); g.replay()` runs one iteration with a single launch.
        Possible because nothing in the step depends on a host value that changes between iterations (the Adam step
        number is a device counter).  The tensors of the keyframe and the parameters must stay where they are; learning
        rates and loss weights are frozen into the graph (re-capture after changing them).  With warmup=True one real
        step is taken first (on the current stream) so that lazily created resources exist before the capture."""
        if warmup:
            self(rs, tile_mask, gt_color, gt_depth, render_mask, gt_semantic)
            torch.cuda.current_stream().synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode="relaxed"):
            self(rs, tile_mask, gt_color, gt_depth, render_mask, gt_semantic)
        return g

    def check(self, auto_resize=False):
        """Reads the device status words and the step counters (one host synchronisation).  Every step whose forward
        overflowed the instance capacity was skipped on the device (parameters, moments and the Adam step number
        untouched) and counted; the count is sticky until this call, so a window can be checked once at its end.  By
        default a non-zero count raises; with auto_resize=True the workspace is re-allocated large enough for what the
        device last reported and the number of skipped steps is returned: repeat them.  Returns the status words
        (list) when nothing was skipped."""
        host = self.status.tolist()
        skipped = int(self.step_state[1].item())
        if skipped:
            self.step_state[1] = 0
            if auto_resize:
                if self.front:
                    back = max(self.back * 2, int(host[_lib.ST_R_BACK] * 1.5) + 65536)
                    self.resize(self.front + back, self.front, back)
                else:
                    self.resize(max(self.capacity * 2, int(host[_lib.ST_NUM_RENDERED] * 1.3) + 4096))
                return skipped
            raise _lib.DqoError("instance capacity exceeded in %d step(s) (capacity %d, front %d, back %d; last R = %d, back "
                                "needs %d): those steps were skipped on the device, repeat them with larger buffers"
                                % (skipped, self.capacity, self.front, self.back, host[_lib.ST_NUM_RENDERED],
                                   host[_lib.ST_R_BACK]))
        return host

    def resize(self, capacity, front_instances=None, back_instances=None):
        """Re-allocates the step workspace for a larger instance capacity (after check() reported skipped steps); the
        parameters, the Adam state and the step counter are untouched, so the skipped steps can simply be repeated.
        Graphs captured before the resize are stale."""
        self.capacity = int(capacity)
        if front_instances is not None:
            self.front, self.back = int(front_instances), int(back_instances or 0)
        if self.front and self.front + self.back > self.capacity:
            raise ValueError("front + back instances exceed the capacity")
        self._alloc()
        self._kf_cache.clear()

    def mark_all_touched(self):
        """Call after writing non-zero values into `self.state` by hand (e.g. restoring a checkpoint)."""
        self.ever.fill_(1)
        self.ever_list.copy_(torch.arange(self.P, dtype=torch.int32, device=self.dev))
        self.ever_count.fill_(self.P)

    def set_binning(self, front_instances, back_instances):
        """Switch between single-phase (0, 0) and two-phase binning; front + back must fit the capacity."""
        if front_instances and front_instances + back_instances > self.capacity:
            raise ValueError("front + back instances exceed the capacity of this step's workspace")
        self.front, self.back = int(front_instances), int(back_instances)

    def rendered(self):
        """Views (no copy) of the colour [3,H,W], depth [1,H,W], depth index [1,H,W] and T [1,H,W] of the last step."""
        import ctypes as C
        outs = [C.c_void_p() for _ in range(4)]
        check(lib().dqo_mapping_step_outputs(self.P, self.M, self.W, self.H, self.capacity, ptr(self.ws),
                                             *[C.byref(o) for o in outs]), "dqo_mapping_step_outputs")
        base = self.ws.data_ptr()
        H, W = self.H, self.W

        def view(o, n, dtype, shape):
            off = o.value - base
            return self.ws[off:off + n * 4].view(dtype).view(shape)

        return (view(outs[0], 3 * H * W, torch.float32, (3, H, W)), view(outs[1], H * W, torch.float32, (1, H, W)),
                view(outs[2], H * W, torch.int32, (1, H, W)), view(outs[3], H * W, torch.float32, (1, H, W)))
