"""Adam step for the parameter arrays either side of ``render``.

``Adam(params, lr)`` has the update rule of ``optax.adam(lr)`` as the reference's
examples/single_image.py:21-40 uses it (b1 0.9, b2 0.999, eps 1e-8, bias-corrected, no weight
decay) and runs each parameter's update as ONE native kernel (csrc/adam.cu) on the current
stream.  ``zero_grad`` drops the gradients (``set_to_none``), so a training step is
``opt.zero_grad(); loss = f(render(...)); loss.backward(); opt.step()``.

``Adam.for_flat(flat, lr)`` is the data-parallel form: ``flat`` is a ``multiview.FlatParams`` (the
five arrays as views of one flat buffer, their gradients as views of the flat buffer the ranks
all-reduce in place), and the step is ONE kernel over all 14*N floats that reads the reduced
gradient buffer where the collective left it -- no unpacking pass between the reduction and the
update.
"""
from __future__ import annotations

import ctypes
from typing import Iterable

import torch

from . import _lib


class Adam:
    def __init__(self, params: Iterable[torch.Tensor], lr: float = 1e-3, b1: float = 0.9,
                 b2: float = 0.999, eps: float = 1e-8, grads: Iterable[torch.Tensor] | None = None):
        """``grads``: fixed gradient buffers, one per parameter, read instead of ``p.grad``"""
        self.params = [p for p in params]
        self.grads = None if grads is None else [g for g in grads]
        if self.grads is not None:
            if len(self.grads) != len(self.params):
                raise ValueError("Adam: one gradient buffer per parameter")
            for p, g in zip(self.params, self.grads):
                if not (g.is_cuda and g.dtype == torch.float32 and g.is_contiguous()
                        and g.numel() == p.numel()):
                    raise TypeError("Adam: gradient buffers must be contiguous float32 CUDA "
                                    "tensors of the parameters' sizes")
        for p in self.params:
            if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                raise TypeError("Adam: parameters must be contiguous float32 CUDA tensors")
        self.lr, self.b1, self.b2, self.eps = float(lr), float(b1), float(b2), float(eps)
        self.m = [torch.zeros_like(p) for p in self.params]
        self.v = [torch.zeros_like(p) for p in self.params]
        self.t = 0

    @classmethod
    def for_flat(cls, flat, lr: float = 1e-3, **kw) -> "Adam":
        """one update kernel over ``flat.flat`` reading ``flat.grad`` (multiview.FlatParams)"""
        return cls([flat.flat], lr=lr, grads=[flat.grad], **kw)

    def zero_grad(self) -> None:
        if self.grads is not None:
            return  # fixed buffers: their owner clears them (FlatParams.zero_grad)
        for p in self.params:
            p.grad = None

    @torch.no_grad()
    def step(self) -> None:
        self.t += 1
        L = _lib.lib()
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        for i, (p, m, v) in enumerate(zip(self.params, self.m, self.v)):
            if self.grads is not None:
                g = self.grads[i]
            elif p.grad is None:
                continue
            else:
                g = p.grad.contiguous()
            rc = L.jsb200_adam_step(stream, p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(),
                                    p.numel(), self.lr, self.b1, self.b2, self.eps, self.t)
            if rc != 0:
                _lib.check_error("adam_step")
