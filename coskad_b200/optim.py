"""Adam over flat buffers: the optimizer half of the training step as one hand-written kernel.

The task modules configure ``torch.optim.Adam(self.parameters(), lr=opt_lr)`` like the reference
(models/hyperbolic_encoder.py:198-217: no weight decay, default betas / eps).  torch's fused multi-tensor kernel spends two
~22 us launches on the encoder's 62 small tensors; with the gradients already in ONE flat bucket (dist.FlatGradBucket) the
parameters and the two moment buffers are made flat as well and the whole update is one pass (``coskad_adam_step``).

``FlatAdam.wrap(opt, bucket)`` returns None -- and the caller keeps ``opt.step()`` -- unless ``opt`` is a plain Adam over exactly
the bucket's float32 CUDA parameters.  The torch optimizer object stays the owner of the hyper-parameters: the learning rate is
read from ``opt.param_groups[0]['lr']`` (a device scalar the schedulers update in place), so ReduceLROnPlateau /
CosineAnnealingLR keep working unchanged.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib


class FlatAdam:
    def __init__(self, opt: torch.optim.Adam, bucket) -> None:
        self.opt, self.bucket = opt, bucket
        g = opt.param_groups[0]
        self.beta1, self.beta2 = (float(b) for b in g['betas'])
        self.eps = float(g['eps'])
        params = bucket.params
        dev = params[0].device
        n = bucket.flat.numel()
        assert n % 4 == 0
        # parameters -> views of one flat buffer, in the bucket's order (same offsets as the gradient views)
        self.flat_p = torch.zeros(n, device=dev, dtype=torch.float32)
        off = 0
        with torch.no_grad():
            for p in params:
                v = self.flat_p[off:off + p.numel()].view_as(p)
                v.copy_(p.data)
                p.data = v
                off += p.numel()
        self.exp_avg = torch.zeros(n, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(n, device=dev, dtype=torch.float32)
        self.step_count = torch.zeros((), device=dev, dtype=torch.int64)
        self._scratch = torch.zeros(2, device=dev, dtype=torch.float32)
        if not torch.is_tensor(g['lr']):
            g['lr'] = torch.tensor(float(g['lr']), dtype=torch.float32, device=dev)
        self._ctx = _lib.context(dev.index if dev.index is not None else torch.cuda.current_device())

    @staticmethod
    def wrap(opt, bucket) -> Optional['FlatAdam']:
        """a FlatAdam for ``opt`` when it is a plain Adam over the attached bucket's parameters, else None"""
        if type(opt) is not torch.optim.Adam or len(opt.param_groups) != 1 or bucket is None or not bucket.attached():
            return None
        g = opt.param_groups[0]
        if g.get('weight_decay', 0) != 0 or g.get('amsgrad', False) or g.get('maximize', False) or g.get('differentiable', False):
            return None
        ps = [p for p in g['params'] if p.requires_grad]
        if len(ps) != len(bucket.params) or any(a is not b for a, b in zip(ps, bucket.params)):
            return None
        if any(p.dtype != torch.float32 or not p.is_cuda for p in ps) or bucket.flat.numel() % 4 != 0:
            return None
        if any(len(opt.state.get(p, {})) for p in ps):            # the torch optimizer has already stepped: keep it
            return None
        return FlatAdam(opt, bucket)

    def intact(self) -> bool:
        """parameters and gradients are still the views this object updates (``module.to()`` / ``zero_grad(set_to_none)`` break them)"""
        off = 0
        for p in self.bucket.params:
            if (p.data.data_ptr() != self.flat_p.data_ptr() + 4 * off or p.grad is None
                    or p.grad.data_ptr() != self.bucket.flat.data_ptr() + 4 * off):
                return False
            off += p.numel()
        return True

    def step(self) -> None:
        lr = self.opt.param_groups[0]['lr']
        if not torch.is_tensor(lr) or lr.device != self.flat_p.device or lr.dtype != torch.float32:
            lr = torch.as_tensor(float(lr), dtype=torch.float32, device=self.flat_p.device)
            self.opt.param_groups[0]['lr'] = lr
        self.opt._opt_called = True               # what lr_scheduler's wrapper of opt.step() records (it warns about call order otherwise)
        c = self._ctx
        c.check(c.lib.coskad_adam_step(c.h, self.flat_p.data_ptr(), self.bucket.flat.data_ptr(), self.exp_avg.data_ptr(),
                                       self.exp_avg_sq.data_ptr(), self.flat_p.numel(), lr.data_ptr(), self.beta1, self.beta2,
                                       self.eps, self.step_count.data_ptr(), self._scratch.data_ptr(),
                                       _lib.stream_ptr(self.flat_p.device)), 'coskad_adam_step')
