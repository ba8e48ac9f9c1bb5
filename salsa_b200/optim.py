"""Optimiser side of the reference's training step, for parameters that live in one flat CUDA
buffer: `Adam` = torch.optim.Adam as `BaseModel.configure_optimizers` builds it (models/interfaces.py:85-95) and
`LearningRateScheduler` = the piecewise-linear learning-rate / beta1 schedule of utilities/learning_utils.py:17-52
(same constructor arguments, `np.interp` over the same step milestones).  The training step that produces the gradients
is salsa_b200.train.SeldTrainer."""
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

    # ---- the same step for a CUDA-graph replayed training step: the scalars of this batch travel through device memory
    def stage_hyper(self):
        """Advances the step count, forms this batch's scalars on the host (crnn_adam_hyper: the same arithmetic as
        crnn_adam_step) and copies them, stream-ordered, into the device buffer `step_staged` reads."""
        if not hasattr(self, '_hyper_host'):
            self._hyper_host = torch.empty(6, dtype=torch.float32).pin_memory()
            self._hyper_dev = torch.zeros(6, dtype=torch.float32, device=self.params.device)
        self.step_count += 1
        _native.check(_native.lib().crnn_adam_hyper(ctypes.c_double(self.lr), ctypes.c_double(self.betas[0]), ctypes.c_double(self.betas[1]),
                                                    ctypes.c_double(self.eps), self.step_count, ctypes.c_void_p(self._hyper_host.data_ptr())))
        self._hyper_dev.copy_(self._hyper_host, non_blocking=True)

    def step_staged(self, flat_grads: torch.Tensor):
        """The update with the scalars staged by `stage_hyper` (one kernel launch, no host state: capturable)."""
        vp = lambda t: ctypes.c_void_p(t.data_ptr())
        with _native.device_of(self.params) as st:
            _native.check(_native.lib().crnn_adam_step_hyper(vp(self.params), vp(flat_grads), vp(self.exp_avg), vp(self.exp_avg_sq),
                                                             self.params.numel(), vp(self._hyper_dev), st))

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
