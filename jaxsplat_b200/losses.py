"""The loss either side of ``render`` (the metric is fwd+bwd of
``loss = mean((img - target) ** 2)``, the loss of the reference's examples/single_image.py:21-30).

``mse_loss_and_grad`` computes the loss and its gradient w.r.t. the image in ONE native kernel
(csrc/loss.cu: image and target read once, cotangent written once) instead of the three or more
elementwise / reduction kernels of an eager framework; the cotangent is what
``((img - target) ** 2).mean().backward()`` hands to ``render``'s backward, bit for bit.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib


def mse_loss_and_grad(img: torch.Tensor, target: torch.Tensor):
    """Returns ``(loss, v_img)``: ``loss`` a float32 device tensor of shape (1,) holding
    ``mean((img - target) ** 2)``, ``v_img = 2 * (img - target) / img.numel()``.  ``target`` may be
    a uint8 image as a data loader delivers it: it is decoded as ``float(t) * (1 / 255)`` inside the
    kernel (bit for bit what ``target.float().mul_(1 / 255)`` stores), never materialised.  Use as
    ``loss, v = mse_loss_and_grad(img.detach(), target); img.backward(gradient=v)``."""
    u8 = target.dtype == torch.uint8
    if not (img.is_cuda and target.is_cuda and img.dtype == torch.float32
            and (u8 or target.dtype == torch.float32) and img.shape == target.shape):
        raise TypeError("mse_loss_and_grad: a float32 CUDA image and a float32 or uint8 CUDA target of one shape")
    img, target = img.contiguous(), target.contiguous()
    n = img.numel()
    v_img = torch.empty_like(img)
    loss = torch.empty(1, dtype=torch.float32, device=img.device)
    fn = _lib.lib().jsb200_mse_loss_grad_u8 if u8 else _lib.lib().jsb200_mse_loss_grad
    rc = fn(
        ctypes.c_void_p(_lib.raw_stream()), ctypes.c_void_p(img.data_ptr()),
        ctypes.c_void_p(target.data_ptr()), n, 2.0 / max(n, 1), ctypes.c_void_p(v_img.data_ptr()),
        ctypes.c_void_p(loss.data_ptr()))
    if rc != 0:
        _lib.check_error("mse_loss_grad")
    return loss, v_img
