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
        self._log_raw: Dict[str, Any] = {}

    #: True when training_step has no Python-side effects beyond in-place updates of device tensors that exist after the
    #: first step (so that Trainer(cuda_graph=True) may replay it); per-epoch state is reset in on_train_epoch_start
    graph_safe = False

    # -- Lightning API used by the reference modules
    def save_hyperparameters(self, *a, **k) -> None:
        pass

    def log(self, name: str, value, *a, **k) -> None:
        # kept as a device scalar: converting here would be one device synchronisation per logged value and step
        self._log_raw[name] = value.detach() if torch.is_tensor(value) else float(value)

    @property
    def _logged(self) -> Dict[str, float]:
        """last logged values as floats (read at epoch granularity; synchronises)"""
        return {k: float(v) for k, v in self._log_raw.items()}

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


_GRAPH_WARMUP = 3


def _flat_tensors(batch) -> bool:
    return isinstance(batch, (list, tuple)) and len(batch) > 0 and all(torch.is_tensor(b) for b in batch)


def _same_layout(batch, static) -> bool:
    return _flat_tensors(batch) and len(batch) == len(static) and \
        all(b.shape == s.shape and b.dtype == s.dtype for b, s in zip(batch, static))


def _make_capturable(opt: torch.optim.Optimizer, device: torch.device) -> None:
    """the optimizer step inside a CUDA graph: step counters on the device, and the learning rate a device scalar the
    schedulers update in place (a Python float would be frozen into the captured kernels' arguments)"""
    for g in opt.param_groups:
        g['capturable'] = True
        if not torch.is_tensor(g['lr']):
            g['lr'] = torch.tensor(float(g['lr']), dtype=torch.float32, device=device)
        if 'initial_lr' in g and torch.is_tensor(g['initial_lr']):
            g['initial_lr'] = float(g['initial_lr'])


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


class TrainStep:
    """One optimisation step: ``training_step`` + backward + flat gradient all-reduce + optimizer.

    Gradients live in one flat bucket (``dist.FlatGradBucket.attach``), so the step has three device phases with stable
    addresses: (A) zero the bucket, forward, backward -- the gradients accumulate into the bucket; (N) one in-place NCCL
    all-reduce of the bucket + the division by the world size; (B) the optimizer.  ``capture`` records A and B as CUDA
    graphs (one graph holding A and B when there is a single rank); ``replay`` enqueues graph A, the collective and graph
    B on the current stream -- three host calls per step and no host synchronisation, so data-parallel runs replay
    exactly like single-GPU ones (the collective itself is not captured: capturing NCCL inside the step graph hung a
    2-GPU replay with torch 2.11 / NCCL 2.28)."""

    def __init__(self, model: 'LightningModule', opt: torch.optim.Optimizer, bucket: 'cdist.FlatGradBucket',
                 device: torch.device) -> None:
        self.model, self.opt, self.bucket, self.device = model, opt, bucket, device
        self.flat_adam = None                      # optim.FlatAdam, created at the first step (the bucket must be attached)
        self._flat_tried = False
        self.graph_a: Optional[torch.cuda.CUDAGraph] = None
        self.graph_b: Optional[torch.cuda.CUDAGraph] = None
        self.static: Optional[List[torch.Tensor]] = None
        self.loss: Optional[torch.Tensor] = None

    def _opt_step(self) -> None:
        """the optimizer: one flat Adam kernel when ``opt`` is a plain Adam over the attached bucket (optim.FlatAdam), else opt.step()"""
        if not self._flat_tried and self.device.type == 'cuda' and not torch.cuda.is_current_stream_capturing():
            self._flat_tried = True
            if not os.environ.get('COSKAD_NO_FLAT_ADAM'):
                from .optim import FlatAdam
                self.flat_adam = FlatAdam.wrap(self.opt, self.bucket)
        if self.flat_adam is not None:
            if not torch.cuda.is_current_stream_capturing() and not self.flat_adam.intact():
                # e.g. module.to() / zero_grad(set_to_none=True) after the first step: the flat buffers no longer alias the
                # parameters -- updating them would silently train nothing
                raise RuntimeError('the parameters / gradients are no longer views of the flat buffers optim.FlatAdam updates; '
                                   'set COSKAD_NO_FLAT_ADAM=1 to keep torch.optim.Adam.step()')
            self.flat_adam.step()
        else:
            self.opt.step()

    def eager(self, batch, batch_idx: int) -> torch.Tensor:
        # a function of its own: no reference to the autograd graph (the loss) survives the step, so a later capture
        # builds fresh AccumulateGrad nodes on the capture stream
        self.bucket.zero_()
        loss = self.model.training_step(batch, batch_idx)
        loss.backward()
        self.bucket.allreduce_()
        self._opt_step()
        return loss.detach()

    @property
    def captured(self) -> bool:
        return self.graph_a is not None

    def capture(self, batch, batch_idx: int) -> None:
        """record the step on static copies of ``batch`` (nothing executes during the capture)"""
        self.static = [b.clone() for b in batch]
        self.bucket.attach()
        torch.cuda.synchronize(self.device)
        split = cdist.world() > 1
        self.graph_a = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_a):
            self.bucket.flat.zero_()
            loss = self.model.training_step(self.static, batch_idx)
            loss.backward()
            if not split:
                self._opt_step()
        self.loss = loss.detach()                  # no reference to the autograd graph is kept
        if split:
            self.graph_b = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_b, pool=self.graph_a.pool()):
                self._opt_step()

    def matches(self, batch) -> bool:
        return self.static is not None and _same_layout(batch, self.static)

    def replay(self, batch) -> torch.Tensor:
        for dst, src in zip(self.static, batch):
            dst.copy_(src, non_blocking=True)
        self.graph_a.replay()
        if self.graph_b is not None:
            self.bucket.allreduce_()               # in place on the bucket the two graphs read and write
            self.graph_b.replay()
        return self.loss


class Trainer:
    def __init__(self, max_epochs: int = 1, device: Optional[torch.device] = None, ckpt_dir: Optional[str] = None,
                 monitor: Optional[str] = None, mode: str = 'max', save_top_k: int = 2, verbose: bool = True,
                 check_val_every_n_epoch: int = 1, cuda_graph: bool = False, callbacks=None, **_ignored) -> None:
        for cb in callbacks or ():          # pytorch_lightning.callbacks.ModelCheckpoint of train_COSKAD.py:70-73
            if hasattr(cb, 'dirpath') and hasattr(cb, 'save_top_k'):
                ckpt_dir = cb.dirpath if cb.dirpath is not None else ckpt_dir
                monitor, mode, save_top_k = cb.monitor or monitor, cb.mode, cb.save_top_k
        self.max_epochs, self.ckpt_dir, self.monitor, self.mode = max_epochs, ckpt_dir, monitor, mode
        self.save_top_k, self.verbose, self.check_val_every_n_epoch = save_top_k, verbose, check_val_every_n_epoch
        self.device = device if device is not None else torch.device('cuda', torch.cuda.current_device())
        # cuda_graph: after _GRAPH_WARMUP eager steps the whole training step (training_step + backward + gradient
        # all-reduce + optimizer) of every full-size batch is one CUDA graph replay (~100 launches and as many tiny
        # allocator / autograd operations per step otherwise bound the step on the host); modules opt in with graph_safe
        self.cuda_graph = cuda_graph
        self.graph_replays = 0
        self.current_epoch = 0
        self.history: List[Dict[str, float]] = []
        self._best: List = []          # heap of (score, path)
        self._data_connector = None
        self.train_dataloader = None

    @classmethod
    def from_argparse_args(cls, args, **kw) -> 'Trainer':
        """pl.Trainer.from_argparse_args(args, default_root_dir=..., max_epochs=..., callbacks=[...]) of train_COSKAD.py:75-78"""
        kw.setdefault('max_epochs', getattr(args, 'ae_epochs', getattr(args, 'max_epochs', 1)))
        kw.setdefault('ckpt_dir', kw.pop('default_root_dir', None))
        return cls(**kw)

    # ------------------------------------------------------------------ fit
    def fit(self, model: LightningModule, train_loader=None, val_loader=None, train_dataloaders=None, val_dataloaders=None):
        train_loader = train_loader if train_loader is not None else train_dataloaders
        val_loader = val_loader if val_loader is not None else val_dataloaders
        model.trainer = self
        if cdist.world() > 1 and not isinstance(train_loader, cdist.ShardedLoader):
            # one process per GPU: every rank trains on its own batches (Lightning's DDP strategy installs a
            # DistributedSampler, train_COSKAD.py:75-85); the center initialisation in setup() walks the same shard and
            # all-reduces its partial sums
            train_loader = cdist.ShardedLoader(train_loader)
        self.train_dataloader = train_loader
        self._data_connector = _DataConnector(train_loader)
        model.to(self.device)
        cdist.broadcast_module_(model)
        model.setup('fit')
        cfg = model.configure_optimizers()
        opt = cfg['optimizer'] if isinstance(cfg, dict) else cfg
        sched = cfg.get('lr_scheduler') if isinstance(cfg, dict) else None
        monitor = (cfg.get('monitor') if isinstance(cfg, dict) else None) or self.monitor
        bucket = cdist.FlatGradBucket(model.parameters()).attach()
        use_graph = self.cuda_graph and bool(getattr(model, 'graph_safe', False))
        if self.cuda_graph and not use_graph and self.verbose and cdist.rank() == 0:
            print(f'cuda_graph: {type(model).__name__} is not graph_safe, training eagerly')
        # always the capturable optimizer form (step counters and learning rate on the device): the eager and the replayed
        # step then run the very same kernels, and with the fixed-order reductions of the training path their loss
        # trajectories are bit-identical
        if self.device.type == 'cuda':
            _make_capturable(opt, self.device)
        step = TrainStep(model, opt, bucket, self.device)
        self.train_step = step
        n_eager = 0
        for epoch in range(self.max_epochs):
            self.current_epoch = epoch
            t0 = time.time()
            model.train()
            model.on_train_epoch_start()
            if hasattr(train_loader, 'set_epoch'):
                train_loader.set_epoch(epoch)
            outputs, nwin = [], 0
            for batch_idx, batch in enumerate(train_loader):
                batch = _to_device(batch, self.device)
                nwin += int(batch[0].shape[0])
                if use_graph and not step.captured and n_eager >= _GRAPH_WARMUP and _flat_tensors(batch) \
                        and int(batch[0].shape[0]) == getattr(train_loader, 'batch_size', int(batch[0].shape[0])):
                    step.capture(batch, batch_idx)
                if step.captured and step.matches(batch):
                    outputs.append(step.replay(batch).clone())
                    self.graph_replays += 1
                    continue
                outputs.append(step.eager(batch, batch_idx))
                n_eager += 1
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


def load_checkpoint(model: nn.Module, path: str, strict: bool = True, trust_pickle: bool = False) -> Dict[str, Any]:
    """checkpoints written here hold tensors, ints and floats only, so they load with ``weights_only=True``;
    ``trust_pickle=True`` opts into unpickling arbitrary objects (a legacy Lightning .ckpt from a trusted source)"""
    ck = torch.load(path, map_location='cpu', weights_only=not trust_pickle)
    sd = ck['state_dict'] if isinstance(ck, dict) and 'state_dict' in ck else ck
    own = model.state_dict()
    # buffers assigned at run time upstream (model.c) may change shape/device; copy what matches by name
    missing = [k for k in own if k not in sd]
    if strict and missing:
        raise KeyError(f'checkpoint {path} lacks keys {missing[:5]}...')
    model.load_state_dict({k: v for k, v in sd.items() if k in own}, strict=False)
    return ck
