"""Drop-in loss modules backed by the sm_100a kernels in libsrst.so.

``StructureTensorLoss``, ``BestBuddyLoss``, ``GramLoss`` and ``PatchwiseStructureTensorLoss`` keep the reference's constructor signatures,
attributes and ``criterion(sr, gt) -> 0-dim fp32 tensor`` contract (reference loss.py:380-413 and
loss.py:78-141; called from train.py:138 and warmup.py:91), so they register through
``config.add_g_criterion(name, module, weight)`` (config.py:122-125) unchanged.  Underneath, each is
a ``torch.autograd.Function`` whose forward/backward enqueue hand-written CUDA kernels on the
current stream through the C ABI of ``include/srst.h``.

No CPU path, no PyTorch fallback: CPU tensors, non-fp32 dtypes and unsupported filter radii raise.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
from torch import nn
from torch.autograd.function import once_differentiable

from . import _cabi, taps as _taps

_WORKSPACES: Dict[Tuple[str, int, int], torch.Tensor] = {}
_DIST_L1 = 0x100  # SRST_BB_DIST_L1 (include/srst.h): OR-ed into the criterion argument, search with dist_norm='l1' (utils.py:166-172)


def _workspace(kind: str, device: torch.device, stream_ptr: int, nbytes: int) -> torch.Tensor:
    """Scratch for one family of entry points (`kind`: "st" or "bb" -- they lay the buffer out differently,
    so they never share one), cached per (kind, device, stream).  Only the 16-byte ticket header has to be
    zero before the first launch; the kernels hand it back zeroed (srst.h, srst_st_workspace_bytes).
    Nothing is cached while a CUDA graph is being captured: an allocation made then belongs to the graph's
    private pool and must not be handed to eager code later."""
    key = (kind, device.index if device.index is not None else torch.cuda.current_device(), stream_ptr)
    ws = _WORKSPACES.get(key)
    if ws is not None and ws.numel() >= nbytes:
        return ws
    ws = torch.zeros(max(nbytes, 4096), dtype=torch.uint8, device=device)
    if not torch.cuda.is_current_stream_capturing():
        _WORKSPACES[key] = ws
    return ws


def _check_pair(x: torch.Tensor, gt: torch.Tensor, who: str) -> None:
    if not (isinstance(x, torch.Tensor) and isinstance(gt, torch.Tensor)):
        raise TypeError(f"{who}: expected two tensors")
    if not (x.is_cuda and gt.is_cuda):
        raise RuntimeError(f"{who}: inputs must be CUDA tensors (this build has no CPU fallback)")
    if x.device != gt.device:
        raise RuntimeError(f"{who}: inputs are on different devices ({x.device} vs {gt.device})")
    if x.dtype != torch.float32 or gt.dtype != torch.float32:
        raise TypeError(f"{who}: inputs must be float32 (got {x.dtype}, {gt.dtype})")
    if x.dim() != 4 or x.shape[1] != 3:
        raise ValueError(f"{who}: expected [B,3,H,W] input, got {tuple(x.shape)}")
    if x.shape != gt.shape:
        raise ValueError(f"{who}: shape mismatch {tuple(x.shape)} vs {tuple(gt.shape)}")


def _no_gt_grad(gt: torch.Tensor, who: str) -> None:
    """The reference would send a gradient into gt through the gathered candidates (loss.py:137-139);
    the training loops never ask for it (gt comes from the data loader) and libsrst has no kernel for it."""
    if gt.requires_grad and torch.is_grad_enabled():
        raise NotImplementedError(f"{who}: gradient w.r.t. gt is not implemented (pass gt.detach())")


def _ptr(t):
    return t.data_ptr() if t is not None else None


def _raw_stream(device: torch.device) -> int:
    """cudaStream_t of torch's current stream on `device` as an integer.  torch.cuda.current_stream() builds a
    Stream object (~5 us); the raw getter behind it is a plain C call."""
    return torch._C._cuda_getCurrentRawStream(device.index if device.index is not None else torch.cuda.current_device())


_SIZES: Dict[Tuple[int, int, int, int, int], Tuple[int, int, int, int]] = {}


def _st_sizes(B: int, H: int, W: int, rs: int = 2, rk: int = 8) -> Tuple[int, int, int, int]:
    """(ds floats padded to 16 bytes, ixy floats, forward workspace bytes, backward workspace bytes) of a
    [B,3,H,W] problem with filter radii (rs, rk), cached per shape.  The last one is 0 for the compiled classes."""
    key = (B, H, W, rs, rk)
    v = _SIZES.get(key)
    if v is None:
        lib = _cabi.lib()
        v = ((B * 3 * H * W + 3) // 4 * 4, lib.srst_st_ixy_floats(B, H, W), lib.srst_st_workspace_bytes_r(B, H, W, rs, rk),
             lib.srst_st_backward_workspace_bytes(B, H, W, rs, rk))
        _SIZES[key] = v
    return v


class _on_device:
    """`with torch.cuda.device(d)` costs ~10 us per call even when d is already current (the normal case in a
    training loop); this enters the context only when the tensor lives on another device."""
    __slots__ = ("_ctx",)

    def __init__(self, device: torch.device):
        self._ctx = None if device.index == torch.cuda.current_device() else torch.cuda.device(device)

    def __enter__(self):
        if self._ctx is not None:
            self._ctx.__enter__()

    def __exit__(self, *exc):
        if self._ctx is not None:
            self._ctx.__exit__(*exc)
        return False


_ST_TAPS: Dict[Tuple[float, float], tuple] = {}


def _st_taps(sigma: float, rho: float, who: str, compiled_only: bool = False):
    """(g*, dg*, r_sigma, k*, r_rho, arrays, generic) as ready-made ctypes arguments, cached per (sigma, rho).
    `generic` is True for radii beyond the compiled shared-memory classes (r_sigma > 4 or r_rho > 12, i.e.
    sigma > 1.1 or rho > 3.1): they run on the library's generic-radius path, which needs scratch planes in the
    workspace and has no fused Pixel / feature variant (`compiled_only`).  Radii above 64 raise."""
    key = (float(sigma), float(rho))
    t = _ST_TAPS.get(key)
    if t is None:
        g, dg = _taps.gaussian_taps(key[0])
        k, _ = _taps.gaussian_taps(key[1])
        rs, rk = len(g) // 2, len(k) // 2
        sup = _cabi.lib().srst_st_supported(rs, rk)
        if not sup:
            raise NotImplementedError(
                f"{who}: filter radii (sigma={sigma} -> {rs}, rho={rho} -> {rk}) exceed what libsrst.so handles (64)")
        t = (_taps.as_c(g), _taps.as_c(dg), rs, _taps.as_c(k), rk, (g, dg, k), sup == 2)  # keep the arrays alive
        _ST_TAPS[key] = t
    if compiled_only and t[6]:
        raise NotImplementedError(
            f"{who}: filter radii (sigma={sigma} -> {t[2]}, rho={rho} -> {t[4]}) are outside the compiled radius classes; "
            "only StructureTensorLoss has a generic-radius path")
    return t


class _StructureTensorLossFn(torch.autograd.Function):
    """autograd boundary of the fused ST loss (the reference lets autograd differentiate ~100
    ATen ops instead; loss.py:399-413).  Saved for backward: per image that needs a gradient, ONE buffer
    holding the ds planes (12 B/px) followed by the Ix, Iy planes (8 B/px); the images themselves are not."""

    @staticmethod
    def forward(ctx, sr, hr, sigma, rho, normalize):
        lib = _cabi.lib()
        sr = sr.contiguous()
        hr = hr.contiguous()
        B, _, H, W = sr.shape
        tp = _st_taps(sigma, rho, "StructureTensorLoss")
        need = ctx.needs_input_grad
        n_ds, n_ixy, ws_bytes, bws_bytes = _st_sizes(B, H, W, tp[2], tp[4])   # n_ds is padded: the ixy part stays 16-byte aligned
        with _on_device(sr.device):
            stream = _raw_stream(sr.device)
            loss = torch.empty((), dtype=torch.float32, device=sr.device)
            saved = [torch.empty(n_ds + n_ixy, dtype=torch.float32, device=sr.device) if n else None for n in need[:2]]
            ds = [t.data_ptr() if t is not None else None for t in saved]
            ixy = [p + 4 * n_ds if p is not None else None for p in ds]
            ws = _workspace("st", sr.device, stream, ws_bytes)
            rc = lib.srst_st_forward(sr.data_ptr(), hr.data_ptr(), B, H, W, tp[0], tp[1], tp[2], tp[3], tp[4],
                                     int(bool(normalize)), 1e-12, loss.data_ptr(), ds[0], ds[1], ixy[0], ixy[1],
                                     ws.data_ptr(), ws.numel(), stream)
        if rc:
            _cabi.check(rc, "srst_st_forward")
        ctx.save_for_backward(saved[0], saved[1])
        ctx.meta = (tp, B, H, W, n_ds, bws_bytes)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        lib = _cabi.lib()
        tp, B, H, W, n_ds, bws_bytes = ctx.meta
        if grad_out.dtype != torch.float32 or not grad_out.is_contiguous():
            grad_out = grad_out.to(torch.float32).contiguous()
        outs = [None, None]
        dev = grad_out.device
        with _on_device(dev):
            stream = _raw_stream(dev)
            for i, saved in enumerate(ctx.saved_tensors):
                if saved is None or not ctx.needs_input_grad[i]:
                    continue
                d_img = torch.empty((B, 3, H, W), dtype=torch.float32, device=dev)
                if bws_bytes:  # generic-radius path: five scratch planes
                    ws = _workspace("st_bwd", dev, stream, bws_bytes)
                    rc = lib.srst_st_backward_ws(saved.data_ptr() + 4 * n_ds, saved.data_ptr(), grad_out.data_ptr(), B, H, W,
                                                 tp[0], tp[1], tp[2], tp[3], tp[4], d_img.data_ptr(), ws.data_ptr(),
                                                 ws.numel(), stream)
                else:
                    rc = lib.srst_st_backward(saved.data_ptr() + 4 * n_ds, saved.data_ptr(), grad_out.data_ptr(), B, H, W,
                                              tp[0], tp[1], tp[2], tp[3], tp[4], d_img.data_ptr(), stream)
                if rc:
                    _cabi.check(rc, "srst_st_backward")
                outs[i] = d_img
        return outs[0], outs[1], None, None, None


class StructureTensorLoss(nn.Module):
    """Structure-tensor loss; same signature and semantics as reference loss.py:380-413.

    ``forward(x, gt)``: x = SR ``[B,3,H,W]``, gt = HR, both fp32 CUDA; returns the mean over all
    pixels of the affine-invariant distance between the det-normalised structure tensors.
    """

    def __init__(self, sigma: float = 0.5, rho: float = 2.0, normalize: bool = True):
        super().__init__()
        self.sigma = sigma
        self.rho = rho
        self.normalize = normalize
        _cabi.lib()  # fail at construction time if the CUDA library is missing

    def st_loss(self, x, gt):
        """Single-sample form (reference loss.py:399-409): x, gt are [3,H,W]."""
        return self.forward(x.unsqueeze(0), gt.unsqueeze(0))

    def forward(self, x, gt):
        _check_pair(x, gt, "StructureTensorLoss")
        return _StructureTensorLossFn.apply(x, gt, self.sigma, self.rho, self.normalize)

    def extra_repr(self) -> str:
        return f"sigma={self.sigma}, rho={self.rho}, normalize={self.normalize}"


def structure_tensor_features(img: torch.Tensor, sigma: float = 0.5, rho: float = 2.0):
    """Diagnostic structure-tensor features of ``img`` ``[B,3,H,W]`` (fp32 CUDA), no gradient: a dict with
    ``J`` ``[B,3,H,W]`` (Jxx, Jyy, Jxy as reference utils.py:212-233 names them), ``eigenvalues`` ``[B,2,H,W]`` (small,
    large), ``orientation`` ``[B,H,W]`` (angle of the dominant-gradient eigenvector against the H axis, radians in
    (-pi/2, pi/2]; the coherent structure runs perpendicular to it) and ``coherence`` ``[B,H,W]``
    (``1 - lambda_small / lambda_large``).  The reference derives orientation / anisotropy only in an exploration
    notebook via the third-party ``structure_tensor`` package; see ``include/srst.h`` (srst_st_features)."""
    if not (isinstance(img, torch.Tensor) and img.is_cuda and img.dtype == torch.float32 and img.dim() == 4
            and img.shape[1] == 3):
        raise TypeError("structure_tensor_features: expected a float32 CUDA tensor [B,3,H,W] (no CPU fallback)")
    lib = _cabi.lib()
    img = img.detach().contiguous()
    B, _, H, W = img.shape
    tp = _st_taps(sigma, rho, "structure_tensor_features", compiled_only=True)
    with _on_device(img.device):
        stream = _raw_stream(img.device)
        J = torch.empty((B, 3, H, W), dtype=torch.float32, device=img.device)
        eig = torch.empty((B, 2, H, W), dtype=torch.float32, device=img.device)
        orient = torch.empty((B, H, W), dtype=torch.float32, device=img.device)
        coher = torch.empty((B, H, W), dtype=torch.float32, device=img.device)
        rc = lib.srst_st_features(img.data_ptr(), B, H, W, tp[0], tp[1], tp[2], tp[3], tp[4], J.data_ptr(), eig.data_ptr(),
                                  orient.data_ptr(), coher.data_ptr(), stream)
    if rc:
        _cabi.check(rc, "srst_st_features")
    return {"J": J, "eigenvalues": eig, "orientation": orient, "coherence": coher}


class _StructureTensorPixelLossFn(torch.autograd.Function):
    """ST loss and the "Pixel" MSE criterion from ONE pass over (sr, gt) per direction.
    Returns the two unweighted terms as one 2-vector [st, mse]; gt carries no gradient."""

    @staticmethod
    def forward(ctx, sr, hr, sigma, rho, normalize):
        lib = _cabi.lib()
        sr = sr.contiguous()
        hr = hr.contiguous()
        B, _, H, W = sr.shape
        tp = _st_taps(sigma, rho, "StructureTensorPixelLoss", compiled_only=True)
        need_sr = ctx.needs_input_grad[0]
        n_ds, n_ixy, ws_bytes, _ = _st_sizes(B, H, W)
        with _on_device(sr.device):
            stream = _raw_stream(sr.device)
            both = torch.empty(2, dtype=torch.float32, device=sr.device)
            saved = torch.empty(n_ds + n_ixy, dtype=torch.float32, device=sr.device) if need_sr else None
            p = saved.data_ptr() if need_sr else None
            ws = _workspace("st", sr.device, stream, ws_bytes)
            rc = lib.srst_stpx_forward(sr.data_ptr(), hr.data_ptr(), B, H, W, tp[0], tp[1], tp[2], tp[3], tp[4],
                                       int(bool(normalize)), 1e-12, both.data_ptr(), p,
                                       p + 4 * n_ds if need_sr else None, ws.data_ptr(), ws.numel(), stream)
        if rc:
            _cabi.check(rc, "srst_stpx_forward")
        ctx.save_for_backward(sr, hr, saved)
        ctx.meta = (tp, n_ds)
        return both

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_both):
        lib = _cabi.lib()
        sr, hr, saved = ctx.saved_tensors
        tp, n_ds = ctx.meta
        if saved is None or not ctx.needs_input_grad[0]:
            return None, None, None, None, None
        B, _, H, W = sr.shape
        with _on_device(sr.device):
            stream = _raw_stream(sr.device)
            if grad_both.dtype != torch.float32 or not grad_both.is_contiguous():
                grad_both = grad_both.to(torch.float32).contiguous()
            d_sr = torch.empty_like(sr)
            g = grad_both.data_ptr()                       # [d/d st, d/d mse]: two adjacent device scalars
            rc = lib.srst_stpx_backward(sr.data_ptr(), hr.data_ptr(), saved.data_ptr() + 4 * n_ds, saved.data_ptr(),
                                        g, g + 4, B, H, W, tp[0], tp[1], tp[2], tp[3], tp[4], d_sr.data_ptr(), stream)
        if rc:
            _cabi.check(rc, "srst_stpx_backward")
        return d_sr, None, None, None, None


class StructureTensorPixelLoss(nn.Module):
    """``st_weight * StructureTensorLoss(sigma, rho, normalize)(x, gt) + pixel_weight * MSELoss()(x, gt)``
    from one fused pass per direction.

    The reference's warm-up / training loops register "Pixel" (``nn.MSELoss``) and "ST" as two criteria
    and evaluate them one after the other on the same two tensors (config.py:71-93, warmup.py:88-96,
    train.py:131-144).  Registering this module once, with weight 1.0, gives the same generator loss
    and gradient; the two unweighted terms of the last call stay on the device in ``last_terms`` for
    logging without a synchronisation."""

    def __init__(self, sigma: float = 0.5, rho: float = 2.0, normalize: bool = True, st_weight: float = 1.0,
                 pixel_weight: float = 1.0):
        super().__init__()
        self.sigma = sigma
        self.rho = rho
        self.normalize = normalize
        self.st_weight = st_weight
        self.pixel_weight = pixel_weight
        self._last = None
        self._w = None   # (device, st_weight, pixel_weight, tensor): the weights as a device 2-vector
        _cabi.lib()

    @property
    def last_terms(self):
        """(st, mse) of the most recent call as two 0-dim device tensors (detached), or None."""
        return None if self._last is None else (self._last[0], self._last[1])

    def forward(self, x, gt):
        _check_pair(x, gt, "StructureTensorPixelLoss")
        _no_gt_grad(gt, "StructureTensorPixelLoss")
        both = _StructureTensorPixelLossFn.apply(x, gt, self.sigma, self.rho, self.normalize)
        self._last = both.detach()
        w = self._w
        if w is None or w[0] != x.device or w[1] != self.st_weight or w[2] != self.pixel_weight:
            w = (x.device, self.st_weight, self.pixel_weight,
                 torch.tensor([self.st_weight, self.pixel_weight], dtype=torch.float32, device=x.device))
            self._w = w
        return torch.dot(both, w[3])

    def extra_repr(self) -> str:
        return (f"sigma={self.sigma}, rho={self.rho}, normalize={self.normalize}, st_weight={self.st_weight}, "
                f"pixel_weight={self.pixel_weight}")


def _patch_backward_gt(mode: int, sr, gt, gt2, gt4, idx, grad_out, taps, criterion: int, stream: int):
    """d loss / d gt of a patch loss (include/srst.h, srst_patch_backward_gt): the reference's gather of the selected
    candidates is differentiable in p2_cat (loss.py:136-139), so a gt that requires grad gets one."""
    lib = _cabi.lib()
    B, _, H, W = sr.shape
    d_gt = torch.empty_like(gt)
    ws = _workspace("bb", sr.device, stream, lib.srst_bb_workspace_bytes(B, H, W))
    if taps is not None:
        g, dg, k = taps
        tp = (_taps.as_c(g), _taps.as_c(dg), len(g) // 2, _taps.as_c(k), len(k) // 2)
    else:
        tp = (None, None, 0, None, 0)
    rc = lib.srst_patch_backward_gt(mode, _ptr(sr), _ptr(gt), _ptr(gt2), _ptr(gt4), _ptr(idx), _ptr(grad_out), B, H, W,
                                    *tp, criterion, _ptr(d_gt), _ptr(ws), ws.numel(), stream)
    _cabi.check(rc, "srst_patch_backward_gt")
    return d_gt


class _BestBuddyLossFn(torch.autograd.Function):
    """autograd boundary of the Best-Buddy loss.  Only the final criterion is differentiable: w.r.t. the SR patches
    and, through the gather of the selected candidates, w.r.t. gt (reference loss.py:135-139); the argmin is not."""

    @staticmethod
    def forward(ctx, sr, gt, alpha, beta, criterion, pyramid, mode="patch"):
        lib = _cabi.lib()
        fwd = lib.srst_bb_forward if mode == "patch" else lib.srst_gram_forward
        sr = sr.contiguous()
        gt = gt.contiguous()
        B, _, H, W = sr.shape
        if H < 12 or W < 12:
            raise ValueError(f"BestBuddyLoss: images must be at least 12x12 (got {H}x{W})")
        with _on_device(sr.device):
            stream = _raw_stream(sr.device)
            if pyramid == "aten":
                # the reference's own op for the HR pyramid (loss.py:123,127)
                with torch.no_grad():
                    gt2 = torch.nn.functional.interpolate(gt, scale_factor=0.5, mode="bicubic",
                                                          align_corners=False).contiguous()
                    gt4 = torch.nn.functional.interpolate(gt, scale_factor=0.25, mode="bicubic",
                                                          align_corners=False).contiguous()
            else:
                gt2 = gt4 = None  # libsrst computes them (srst_bb_pyramid taps)
            N = (H // 3) * (W // 3)
            idx = torch.empty((B, N), dtype=torch.int64, device=sr.device)
            loss = torch.empty((), dtype=torch.float32, device=sr.device)
            nbytes = lib.srst_bb_workspace_bytes(B, H, W)
            ws = _workspace("bb", sr.device, stream, nbytes)
            rc = fwd(_ptr(sr), _ptr(gt), _ptr(gt2), _ptr(gt4), B, H, W, float(alpha), float(beta),
                                     int(criterion), _ptr(idx), _ptr(loss), _ptr(ws), ws.numel(),
                                     stream)
        _cabi.check(rc, "srst_bb_forward" if mode == "patch" else "srst_gram_forward")
        ctx.save_for_backward(sr, gt, gt2, gt4, idx)
        ctx.criterion = int(criterion)
        ctx.mode = mode
        ctx.mark_non_differentiable(idx)
        return loss, idx

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out, _grad_idx):
        lib = _cabi.lib()
        sr, gt, gt2, gt4, idx = ctx.saved_tensors
        B, _, H, W = sr.shape
        bwd = lib.srst_bb_backward if ctx.mode == "patch" else lib.srst_gram_backward
        if not (ctx.needs_input_grad[0] or ctx.needs_input_grad[1]):
            return None, None, None, None, None, None, None
        grad_out = grad_out.to(torch.float32).contiguous()
        d_sr = d_gt = None
        with _on_device(sr.device):
            stream = _raw_stream(sr.device)
            if ctx.needs_input_grad[0]:
                d_sr = torch.empty_like(sr)
                nbytes = lib.srst_bb_workspace_bytes(B, H, W)
                ws = _workspace("bb", sr.device, stream, nbytes)
                rc = bwd(_ptr(sr), _ptr(gt), _ptr(gt2), _ptr(gt4), _ptr(idx), _ptr(grad_out), B, H, W,
                         ctx.criterion, _ptr(d_sr), _ptr(ws), ws.numel(), stream)
                _cabi.check(rc, "srst_bb_backward" if ctx.mode == "patch" else "srst_gram_backward")
            if ctx.needs_input_grad[1]:
                d_gt = _patch_backward_gt(0 if ctx.mode == "patch" else 1, sr, gt, gt2, gt4, idx, grad_out, None,
                                          ctx.criterion, stream)
        return d_sr, d_gt, None, None, None, None, None


class _BestBuddyGeometryFn(torch.autograd.Function):
    """BestBuddyLoss with a non-default (ksize, pad, stride): the exact all-pairs path of libsrst (include/srst.h,
    srst_bbg_forward / srst_bbg_backward).  Differentiable w.r.t. the SR patches (folded back over overlapping patches)
    and, through the gather of the selected candidates, w.r.t. gt (loss.py:135-139); the argmin is not."""

    @staticmethod
    def forward(ctx, sr, gt, alpha, beta, criterion, pyramid, geom):
        lib = _cabi.lib()
        sr = sr.contiguous()
        gt = gt.contiguous()
        B, _, H, W = sr.shape
        ks, pad, st = geom
        N = lib.srst_bbg_num_patches(H, W, ks, pad, st)
        nbytes = lib.srst_bbg_workspace_bytes(B, H, W, ks, pad, st)
        if N <= 0 or nbytes == 0:
            raise ValueError(f"BestBuddyLoss: a {H}x{W} image has a pyramid level without a single "
                             f"ksize={ks}, pad={pad}, stride={st} patch (F.unfold would raise too)")
        with _on_device(sr.device):
            stream = _raw_stream(sr.device)
            if pyramid == "aten":
                with torch.no_grad():
                    gt2 = torch.nn.functional.interpolate(gt, scale_factor=0.5, mode="bicubic",
                                                          align_corners=False).contiguous()
                    gt4 = torch.nn.functional.interpolate(gt, scale_factor=0.25, mode="bicubic",
                                                          align_corners=False).contiguous()
            else:
                gt2 = gt4 = None
            idx = torch.empty((B, N), dtype=torch.int64, device=sr.device)
            loss = torch.empty((), dtype=torch.float32, device=sr.device)
            ws = _workspace("bbg", sr.device, stream, nbytes)
            rc = lib.srst_bbg_forward(_ptr(sr), _ptr(gt), _ptr(gt2), _ptr(gt4), B, H, W, ks, pad, st, float(alpha),
                                      float(beta), int(criterion), _ptr(idx), _ptr(loss), _ptr(ws), ws.numel(), stream)
        _cabi.check(rc, "srst_bbg_forward")
        ctx.save_for_backward(sr, gt, gt2, gt4, idx)
        ctx.criterion = int(criterion)
        ctx.geom = geom
        ctx.mark_non_differentiable(idx)
        return loss, idx

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out, _grad_idx):
        lib = _cabi.lib()
        sr, gt, gt2, gt4, idx = ctx.saved_tensors
        if not (ctx.needs_input_grad[0] or ctx.needs_input_grad[1]):
            return None, None, None, None, None, None, None
        B, _, H, W = sr.shape
        ks, pad, st = ctx.geom
        grad_out = grad_out.to(torch.float32).contiguous()
        with _on_device(sr.device):
            stream = _raw_stream(sr.device)
            d_sr = torch.empty_like(sr) if ctx.needs_input_grad[0] else None
            d_gt = torch.empty_like(gt) if ctx.needs_input_grad[1] else None
            ws = _workspace("bbg", sr.device, stream, lib.srst_bbg_workspace_bytes(B, H, W, ks, pad, st))
            rc = lib.srst_bbg_backward(_ptr(sr), _ptr(gt), _ptr(gt2), _ptr(gt4), _ptr(idx), _ptr(grad_out), B, H, W,
                                       ks, pad, st, ctx.criterion, _ptr(d_sr), _ptr(d_gt), _ptr(ws), ws.numel(), stream)
        _cabi.check(rc, "srst_bbg_backward")
        return d_sr, d_gt, None, None, None, None, None


class BestBuddyLoss(nn.Module):
    """Best-Buddy loss; same signature and semantics as reference loss.py:78-141.

    ``forward(x, gt)``: for every 3x3 SR patch pick the HR candidate patch (three pyramid levels)
    minimising ``alpha*|sr-cand|^2 + beta*|gt-cand|^2`` and return the L1 (or MSE) distance to it.
    The reference's default patch geometry (ksize=3, pad=0, stride=3) runs the tuned filter + exact re-scoring search;
    any other ``ksize`` (1..8), ``pad`` and ``stride`` -- overlapping patches, gaps, zero padding (F.unfold semantics,
    loss.py:116-129) -- runs an exact all-pairs search kernel.  ``dist_norm='l1'`` (utils.py:166-172) scores every pair
    exactly with
    ``alpha*sum|sr-cand| + beta*sum|gt-cand|`` (the reference materialises a [B,N,M,27] tensor for it).

    ``pyramid``: "fused" (default) lets libsrst build the two HR pyramid levels with the cubic taps of
    ``F.interpolate(mode='bicubic', align_corners=False)`` (equal to 2e-7; 0.03 ms instead of 0.43 ms at
    batch 64 x 192x192); "aten" calls the reference's own op.
    ``last_indices`` holds the argmin indices ``[B,N]`` (int64) of the most recent call.
    """

    def __init__(self, alpha: float = 1.0, beta: float = 1.0, ksize: int = 3, pad: int = 0, stride: int = 3,
                 dist_norm: str = "l2", criterion: str = "l1", pyramid: str = "fused"):
        super().__init__()
        self.alpha = alpha
        self.beta = beta
        self.ksize = ksize
        self.pad = pad
        self.stride = stride
        self.dist_norm = dist_norm
        if criterion == "l1":
            self.criterion = torch.nn.L1Loss()   # kept for repr/config parity (config.py:133-139)
            self._crit = 0
        elif criterion == "l2" or criterion == "mse":
            self.criterion = torch.nn.MSELoss()
            self._crit = 1
        else:
            raise NotImplementedError("%s criterion has not been implmented." % criterion)  # loss.py:113
        if dist_norm not in ("l1", "l2"):
            raise NotImplementedError("%s norm has not been supported." % dist_norm)        # utils.py:189
        self._geom = None   # None: the tuned (3, 0, 3) kernels; else the generic-geometry path
        if (ksize, pad, stride) != (3, 0, 3):
            if not all(isinstance(v, int) and not isinstance(v, bool) for v in (ksize, pad, stride)):
                raise TypeError("BestBuddyLoss: ksize, pad and stride must be ints")
            if not _cabi.lib().srst_bbg_supported(ksize, pad, stride):
                raise NotImplementedError(
                    f"BestBuddyLoss: libsrst.so has kernels for 1 <= ksize <= 8, pad >= 0, stride >= 1 "
                    f"(got ksize={ksize}, pad={pad}, stride={stride})")
            self._geom = (ksize, pad, stride)
        if dist_norm == "l1":
            self._crit |= _DIST_L1
        if pyramid not in ("aten", "fused"):
            raise ValueError("pyramid must be 'aten' or 'fused'")
        self.pyramid = pyramid
        self.last_indices = None
        _cabi.lib()

    def forward(self, x, gt):
        _check_pair(x, gt, "BestBuddyLoss")
        if self._geom is not None:
            loss, idx = _BestBuddyGeometryFn.apply(x, gt, self.alpha, self.beta, self._crit, self.pyramid, self._geom)
        else:
            loss, idx = _BestBuddyLossFn.apply(x, gt, self.alpha, self.beta, self._crit, self.pyramid)
        self.last_indices = idx
        return loss


class GramLoss(nn.Module):
    """Gram loss; same signature and semantics as reference loss.py:146-225: the best-buddy search
    and the final criterion run on the 3x3 Gram matrix of every 3x3x3 patch instead of its pixels.
    Only ``ksize=3`` has kernels (the reference's default); ``dist_norm`` may be 'l2' or 'l1'."""

    def __init__(self, alpha: float = 1.0, beta: float = 1.0, ksize: int = 3, dist_norm: str = "l2",
                 criterion: str = "l1", pyramid: str = "fused"):
        super().__init__()
        self.alpha = alpha
        self.beta = beta
        self.ksize = ksize
        self.dist_norm = dist_norm
        if criterion == "l1":
            self.criterion = torch.nn.L1Loss()
            self._crit = 0
        elif criterion == "l2" or criterion == "mse":
            self.criterion = torch.nn.MSELoss()
            self._crit = 1
        else:
            raise NotImplementedError("%s criterion has not been implmented." % criterion)  # loss.py:178
        if dist_norm not in ("l1", "l2"):
            raise NotImplementedError("%s norm has not been supported." % dist_norm)        # utils.py:189
        if ksize != 3:
            raise NotImplementedError("GramLoss: libsrst.so implements ksize=3 only")
        if dist_norm == "l1":
            self._crit |= _DIST_L1
        if pyramid not in ("aten", "fused"):
            raise ValueError("pyramid must be 'aten' or 'fused'")
        self.pyramid = pyramid
        self.last_indices = None
        _cabi.lib()

    def forward(self, x, gt):
        _check_pair(x, gt, "GramLoss")
        loss, idx = _BestBuddyLossFn.apply(x, gt, self.alpha, self.beta, self._crit, self.pyramid, "gram")
        self.last_indices = idx
        return loss


class _PatchwiseStLossFn(torch.autograd.Function):
    """autograd boundary of the patchwise structure-tensor loss: differentiable through the final
    criterion and the descriptors of the SR patch and of the selected candidate (argmin is not; loss.py:366-371)."""

    @staticmethod
    def forward(ctx, sr, gt, sigma, rho, alpha, beta, criterion, pyramid):
        lib = _cabi.lib()
        sr = sr.contiguous()
        gt = gt.contiguous()
        B, _, H, W = sr.shape
        if H < 12 or W < 12:
            raise ValueError(f"PatchwiseStructureTensorLoss: images must be at least 12x12 (got {H}x{W})")
        g, dg = _taps.gaussian_taps(float(sigma))
        k, _ = _taps.gaussian_taps(float(rho))
        with _on_device(sr.device):
            stream = _raw_stream(sr.device)
            if pyramid == "aten":
                with torch.no_grad():  # the reference's own op for the HR pyramid (loss.py:353,356)
                    gt2 = torch.nn.functional.interpolate(gt, scale_factor=0.5, mode="bicubic",
                                                          align_corners=False).contiguous()
                    gt4 = torch.nn.functional.interpolate(gt, scale_factor=0.25, mode="bicubic",
                                                          align_corners=False).contiguous()
            else:
                gt2 = gt4 = None
            N = (H // 3) * (W // 3)
            idx = torch.empty((B, N), dtype=torch.int64, device=sr.device)
            loss = torch.empty((), dtype=torch.float32, device=sr.device)
            ws = _workspace("bb", sr.device, stream, lib.srst_bb_workspace_bytes(B, H, W))
            rc = lib.srst_pst_forward(_ptr(sr), _ptr(gt), _ptr(gt2), _ptr(gt4), B, H, W, _taps.as_c(g), _taps.as_c(dg),
                                      len(g) // 2, _taps.as_c(k), len(k) // 2, float(alpha), float(beta),
                                      int(criterion), _ptr(idx), _ptr(loss), _ptr(ws), ws.numel(),
                                      stream)
        _cabi.check(rc, "srst_pst_forward")
        ctx.save_for_backward(sr, gt, gt2, gt4, idx)
        ctx.taps = (g, dg, k)
        ctx.criterion = int(criterion)
        ctx.mark_non_differentiable(idx)
        return loss, idx

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out, _grad_idx):
        lib = _cabi.lib()
        sr, gt, gt2, gt4, idx = ctx.saved_tensors
        g, dg, k = ctx.taps
        B, _, H, W = sr.shape
        if not (ctx.needs_input_grad[0] or ctx.needs_input_grad[1]):
            return (None,) * 8
        grad_out = grad_out.to(torch.float32).contiguous()
        d_sr = d_gt = None
        with _on_device(sr.device):
            stream = _raw_stream(sr.device)
            if ctx.needs_input_grad[0]:
                d_sr = torch.empty_like(sr)
                ws = _workspace("bb", sr.device, stream, lib.srst_bb_workspace_bytes(B, H, W))
                rc = lib.srst_pst_backward(_ptr(sr), _ptr(gt), _ptr(gt2), _ptr(gt4), _ptr(idx), _ptr(grad_out), B, H, W,
                                           _taps.as_c(g), _taps.as_c(dg), len(g) // 2, _taps.as_c(k), len(k) // 2,
                                           ctx.criterion, _ptr(d_sr), _ptr(ws), ws.numel(), stream)
                _cabi.check(rc, "srst_pst_backward")
            if ctx.needs_input_grad[1]:
                d_gt = _patch_backward_gt(2, sr, gt, gt2, gt4, idx, grad_out, (g, dg, k), ctx.criterion, stream)
        return (d_sr, d_gt) + (None,) * 6


class PatchwiseStructureTensorLoss(nn.Module):
    """Patchwise structure-tensor loss; same signature and semantics as reference loss.py:292-375:
    the best-buddy search and the final criterion run on the det-normalised structure tensor of
    every 3x3 patch (seen as a 3x3 image, zero 'same' padding) instead of its pixels.
    Only ``ksize=3`` has kernels (the reference's default); ``dist_norm`` may be 'l2' or 'l1'."""

    def __init__(self, sigma: float = 0.5, rho: float = 2, alpha: float = 1.0, beta: float = 1.0, ksize: int = 3,
                 dist_norm: str = "l2", criterion: str = "l1", pyramid: str = "fused"):
        super().__init__()
        self.alpha = alpha
        self.beta = beta
        self.ksize = ksize
        self.dist_norm = dist_norm
        self.sigma = sigma
        self.rho = rho
        if criterion == "l1":
            self.criterion = torch.nn.L1Loss(reduction="mean")
            self._crit = 0
        elif criterion == "l2" or criterion == "mse":
            self.criterion = torch.nn.MSELoss(reduction="mean")
            self._crit = 1
        else:
            raise NotImplementedError("%s criterion has not been supported." % criterion)  # loss.py:323
        if dist_norm not in ("l1", "l2"):
            raise NotImplementedError("%s norm has not been supported." % dist_norm)       # utils.py:189
        if ksize != 3:
            raise NotImplementedError("PatchwiseStructureTensorLoss: libsrst.so implements ksize=3 only")
        if dist_norm == "l1":
            self._crit |= _DIST_L1
        if pyramid not in ("aten", "fused"):
            raise ValueError("pyramid must be 'aten' or 'fused'")
        self.pyramid = pyramid
        self.last_indices = None
        _cabi.lib()

    def forward(self, x, gt):
        _check_pair(x, gt, "PatchwiseStructureTensorLoss")
        loss, idx = _PatchwiseStLossFn.apply(x, gt, self.sigma, self.rho, self.alpha, self.beta, self._crit,
                                             self.pyramid)
        self.last_indices = idx
        return loss
