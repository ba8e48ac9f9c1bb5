"""Optimiser side of the reference's training step, for parameters that live in one flat CUDA
buffer: `Adam` = torch.optim.Adam as `BaseModel.configure_optimizers` builds it (models/interfaces.py:85-95) and
`LearningRateScheduler` = the piecewise-linear learning-rate / beta1 schedule of utilities/learning_utils.py:17-52
(same constructor arguments, `np.interp` over the same step milestones).  The backward pass that would produce the
gradients is not part of this package yet."""
import ctypes

import numpy as np
import torch

from . import _native

__all__ = ['Adam', 'LearningRateScheduler']


class LearningRateScheduler:
    def __init__(self, steps_per_epoch, max_epochs: int = 50, milestones=(0, 0.45, 0.9, 1.0), lrs=(1e-4, 1e-2, 1e-3, 1e-4),
                 moms=(0.9, 0.8, 0.9, 0.9)):
        self.steps_per_epoch, self.max_epochs = steps_per_epoch, max_epochs
        self.milestones, self.lrs, self.moms = milestones, lrs, moms
        self.n_steps = int(self.max_epochs * self.steps_per_epoch)
        self.step_milestones = [int(i * self.n_steps) for i in self.milestones]

    def at(self, current_epoch: int, batch_idx: int):
        """(lr, beta1) `on_train_batch_start` sets for this batch (learning_utils.py:44-52)."""
        step = current_epoch * self.steps_per_epoch + batch_idx
        return float(np.interp(step, self.step_milestones, self.lrs)), float(np.interp(step, self.step_milestones, self.moms))

    def apply(self, optimizer, current_epoch: int, batch_idx: int):
        optimizer.lr, mom = self.at(current_epoch, batch_idx)
        optimizer.betas = (mom, optimizer.betas[1])


class Adam:
    """Adam over ONE flat float32 CUDA tensor of parameters (views of it are the model's tensors)."""

    def __init__(self, flat_params: torch.Tensor, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        if not (flat_params.is_cuda and flat_params.dtype == torch.float32 and flat_params.is_contiguous()):
            raise ValueError('flat_params must be a contiguous float32 CUDA tensor')
        self.params = flat_params
        self.lr, self.betas, self.eps = lr, tuple(betas), eps
        self.exp_avg = torch.zeros_like(flat_params)
        self.exp_avg_sq = torch.zeros_like(flat_params)
        self.step_count = 0

    def step(self, flat_grads: torch.Tensor):
        if flat_grads.shape != self.params.shape or flat_grads.dtype != torch.float32 or not flat_grads.is_cuda:
            raise ValueError('flat_grads must match the parameter buffer')
        self.step_count += 1
        vp = lambda t: ctypes.c_void_p(t.data_ptr())
        _native.same_device(self.params, flat_grads)
        with _native.device_of(self.params) as st:
            _native.check(_native.lib().crnn_adam_step(vp(self.params), vp(flat_grads.contiguous()), vp(self.exp_avg), vp(self.exp_avg_sq),
                                                       self.params.numel(), ctypes.c_double(self.lr), ctypes.c_double(self.betas[0]),
                                                       ctypes.c_double(self.betas[1]), ctypes.c_double(self.eps), self.step_count, st))
