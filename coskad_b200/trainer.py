"""Minimal in-repo replacement of the pytorch_lightning subset COSKAD uses (train_COSKAD.py:65-85,
eval_COSKAD.py:115-116): ``LightningModule`` hooks, ``Trainer.fit / predict``, top-k ``ModelCheckpoint``.

pytorch_lightning is not installable here and is orchestration, not hot path.  Data
parallelism is one process per GPU (torchrun); gradients go through one flat NCCL all-reduce per step
(coskad_b200.dist.FlatGradBucket) instead of Lightning's DDP wrapper.  Checkpoints are ``torch.save``d
dicts with a ``state_dict`` whose keys equal the reference's (``model.encoder.model.0.gcn.A`` ...).
"""
from __future__ import annotations

import heapq
import os
import time
from typing import Any, Dict, List, Optional

import torch
import torch.nn as nn

from . import dist as cdist


class LightningModule(nn.Module):
    """the hook surface the COSKAD task modules rely on ("Trainer hooks")"""

    def __init__(self) -> None:
        super().__init__()
        self.trainer: Optional['Trainer'] = None
        self._logged: Dict[str, float] = {}

    # -- Lightning API used by the reference modules
    def save_hyperparameters(self, *a, **k) -> None:
        pass

    def log(self, name: str, value, *a, **k) -> None:
        self._logged[name] = float(value.detach()) if torch.is_tensor(value) else float(value)

    @property
    def device(self) -> torch.device:
        try:
            return next(self.parameters()).device
        except StopIteration:
            return torch.device('cpu')

    def setup(self, stage: Optional[str] = None) -> None: ...
    def on_train_epoch_start(self) -> None: ...
    def training_epoch_end(self, outputs) -> None: ...
    def on_train_epoch_end(self) -> None: ...
    def on_validation_start(self) -> None: ...
    def validation_step(self, batch, batch_idx): ...
    def validation_epoch_end(self, outputs): ...
    def configure_optimizers(self): raise NotImplementedError


def _to_device(batch, device):
    if torch.is_tensor(batch):
        return batch.to(device, non_blocking=True)
    if isinstance(batch, (list, tuple)):
        return [_to_device(b, device) for b in batch]
    return batch


class _LoaderSource:                      # self.trainer._data_connector._train_dataloader_source.dataloader()
    def __init__(self, loader): self._loader = loader
    def dataloader(self): return self._loader


class _DataConnector:
    def __init__(self, loader): self._train_dataloader_source = _LoaderSource(loader)


class Trainer:
    def __init__(self, max_epochs: int = 1, device: Optional[torch.device] = None, ckpt_dir: Optional[str] = None,
                 monitor: Optional[str] = None, mode: str = 'max', save_top_k: int = 2, verbose: bool = True,
                 check_val_every_n_epoch: int = 1, **_ignored) -> None:
        self.max_epochs, self.ckpt_dir, self.monitor, self.mode = max_epochs, ckpt_dir, monitor, mode
        self.save_top_k, self.verbose, self.check_val_every_n_epoch = save_top_k, verbose, check_val_every_n_epoch
        self.device = device if device is not None else torch.device('cuda', torch.cuda.current_device())
        self.current_epoch = 0
        self.history: List[Dict[str, float]] = []
        self._best: List = []          # heap of (score, path)
        self._data_connector = None
        self.train_dataloader = None

    # ------------------------------------------------------------------ fit
    def fit(self, model: LightningModule, train_loader=None, val_loader=None, train_dataloaders=None, val_dataloaders=None):
        train_loader = train_loader if train_loader is not None else train_dataloaders
        val_loader = val_loader if val_loader is not None else val_dataloaders
        model.trainer = self
        self.train_dataloader = train_loader
        self._data_connector = _DataConnector(train_loader)
        model.to(self.device)
        cdist.broadcast_module_(model)
        model.setup('fit')
        cfg = model.configure_optimizers()
        opt = cfg['optimizer'] if isinstance(cfg, dict) else cfg
        sched = cfg.get('lr_scheduler') if isinstance(cfg, dict) else None
        monitor = (cfg.get('monitor') if isinstance(cfg, dict) else None) or self.monitor
        bucket = cdist.FlatGradBucket(model.parameters())
        for epoch in range(self.max_epochs):
            self.current_epoch = epoch
            t0 = time.time()
            model.train()
            model.on_train_epoch_start()
            outputs, nwin = [], 0
            for batch_idx, batch in enumerate(train_loader):
                batch = _to_device(batch, self.device)
                loss = model.training_step(batch, batch_idx)
                opt.zero_grad(set_to_none=True)
                loss.backward()
                bucket.allreduce_()
                opt.step()
                outputs.append(loss.detach())
                nwin += int(batch[0].shape[0])
            model.training_epoch_end(outputs)
            model.on_train_epoch_end()
            logs = dict(model._logged)
            logs['epoch'] = epoch
            logs['train_loss_mean'] = float(torch.stack(outputs).mean()) if outputs else float('nan')
            if val_loader is not None and (epoch + 1) % self.check_val_every_n_epoch == 0:
                model.eval()
                model.on_validation_start()
                with torch.no_grad():
                    outs = [model.validation_step(_to_device(b, self.device), i) for i, b in enumerate(val_loader)]
                    model.validation_epoch_end(outs)
                logs.update(model._logged)
            if sched is not None:
                if isinstance(sched, torch.optim.lr_scheduler.ReduceLROnPlateau):
                    if monitor in logs:
                        sched.step(logs[monitor])
                else:
                    sched.step()
            torch.cuda.synchronize(self.device)
            logs['epoch_seconds'] = time.time() - t0
            logs['train_windows_per_s'] = nwin * cdist.world() / max(logs['epoch_seconds'], 1e-9)
            self.history.append(logs)
            if self.verbose and cdist.rank() == 0:
                print('epoch', epoch, {k: (round(v, 6) if isinstance(v, float) else v) for k, v in logs.items()})
            self._checkpoint(model, logs, monitor)
        return self

    def _checkpoint(self, model, logs, monitor) -> None:
        """ModelCheckpoint(save_top_k=2, monitor='validation_auc' | 'loss') of train_COSKAD.py:70-73"""
        if self.ckpt_dir is None or cdist.rank() != 0:
            return
        key = monitor if monitor in logs else ('loss' if 'loss' in logs else 'train_loss_mean')
        score = logs[key] if (self.mode == 'max' and key == monitor) else -logs[key]
        os.makedirs(self.ckpt_dir, exist_ok=True)
        path = os.path.join(self.ckpt_dir, f'epoch={logs["epoch"]}-{key}={logs[key]:.6f}.ckpt')
        if len(self._best) < self.save_top_k or score > self._best[0][0]:
            torch.save({'state_dict': model.state_dict(), 'epoch': logs['epoch'], key: logs[key]}, path)
            heapq.heappush(self._best, (score, path))
            while len(self._best) > self.save_top_k:
                _, old = heapq.heappop(self._best)
                if os.path.exists(old):
                    os.remove(old)

    @property
    def best_checkpoints(self) -> List[str]:
        return [p for _, p in sorted(self._best, reverse=True)]

    # ------------------------------------------------------------------ predict
    @torch.no_grad()
    def predict(self, model: LightningModule, dataloaders=None, ckpt_path: Optional[str] = None,
                return_predictions: bool = True) -> List[Any]:
        """model.forward over the loader in eval mode.  Under torchrun every rank takes the batches
        ``i % world == rank`` and the per-batch outputs are all-gathered back in loader order (the reference's
        ``strategy='ddp'`` predict keeps only the local shard, eval_COSKAD.py:115-116 -- fixed here)."""
        model.trainer = self
        if ckpt_path:
            load_checkpoint(model, ckpt_path)
        model.to(self.device).eval()
        w, r = cdist.world(), cdist.rank()
        local = []
        for i, batch in enumerate(dataloaders):
            if i % w == r:
                out = model(_to_device(batch, self.device))
                local.append((i, tuple(o.detach().cpu() if torch.is_tensor(o) else o for o in out)))
        if w > 1:
            import torch.distributed as dist
            gathered = [None] * w
            dist.all_gather_object(gathered, local)
            local = sorted((it for part in gathered for it in part), key=lambda t: t[0])
        return [o for _, o in local]


def load_checkpoint(model: nn.Module, path: str, strict: bool = True) -> Dict[str, Any]:
    ck = torch.load(path, map_location='cpu', weights_only=False)
    sd = ck['state_dict'] if isinstance(ck, dict) and 'state_dict' in ck else ck
    own = model.state_dict()
    # buffers assigned at run time upstream (model.c) may change shape/device; copy what matches by name
    missing = [k for k in own if k not in sd]
    if strict and missing:
        raise KeyError(f'checkpoint {path} lacks keys {missing[:5]}...')
    model.load_state_dict({k: v for k, v in sd.items() if k in own}, strict=False)
    return ck
