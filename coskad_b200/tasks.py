"""Task modules: the COSKAD LightningModules re-hosted on the CUDA hot path.

    LitEncoder      models/hyperbolic_encoder.py:42-305 (hyperbolic) and
                    models/euclidean_encoder_staticCenter.py / _dynamicCenter.py (Euclidean, flag static_center)
    LitAutoEncoder  models/euclidean_autoencoder.py
Same constructor (``args`` namespace), hooks, logged names and config flags (use_decoder, use_vae,
hyperbolic, static_center).  Differences, all documented in DESIGN.md:
  * the center is the finalisation of float64 partial sums that are ALL-REDUCED across ranks (upstream each
    DDP rank keeps its own unsynchronised center, hyperbolic_encoder.py:149-155,175-183) and the latents are
    never concatenated (upstream ``cumt = torch.cat(...)`` grows O(N));
  * validation scoring / aggregation is one batched device pass (aggregate.score_and_aggregate) instead of
    the Python triple loop with one D2H sync per window.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F
from torch.optim import Adam

from . import _lib, aggregate, dist as cdist, gmath
from .losses import calc_reg_loss
from .sts import STSAE, STSE
from .trainer import LightningModule


def _joints(args) -> int:
    if getattr(args, 'dataset_headless', False):
        return 14
    if getattr(args, 'dataset_kp18_format', False):
        return 18
    return 17


def light_processing_data(data):
    """utils/model_utils.py:108-147: list of per-batch output tuples -> concatenated numpy arrays"""
    cat = lambda idx: np.concatenate([np.asarray(d[idx].cpu() if torch.is_tensor(d[idx]) else d[idx]) for d in data], axis=0)
    n = len(data[0])
    if n == 6:
        return cat(0), cat(1), cat(2), cat(3), cat(4), cat(5)
    if n == 4:
        return cat(0), cat(1), cat(2), cat(3)
    return cat(0), cat(2), cat(3), cat(4)


def load_gt_table(args) -> Tuple[List[Tuple[int, int, int]], Dict[Tuple[int, int], np.ndarray]]:
    """(clips, gts): from ``args.gt_table`` (in-memory, synthetic runs) or the ``{scene}_{clip}.npy`` files of
    ``args.gt_path`` in sorted order (eval_COSKAD.py:126-128)."""
    if getattr(args, 'gt_table', None) is not None:
        return args.gt_table
    all_gts = sorted(fn for fn in os.listdir(args.gt_path) if fn.endswith('.npy'))
    clips, gts = [], {}
    for fn in all_gts:
        scene, clip = int(fn.split('_')[0]), int(fn.split('_')[1].split('.')[0])
        gt = np.load(os.path.join(args.gt_path, fn))
        clips.append((scene, clip, int(gt.shape[0])))
        gts[(scene, clip)] = gt
    return clips, gts


def hr_ubnormal(path_to_boolean_masks: str) -> Dict[Tuple[int, int], np.ndarray]:
    """utils/model_utils.py:149-161: ``{scene}_{clip}.npy`` boolean frame masks of the UBnormal HR subset, keyed (scene, clip)"""
    from glob import glob
    masks = {}
    for path in glob(path_to_boolean_masks):
        scene_id, clip_id = map(int, os.path.basename(path).split('.')[0].split('_'))
        masks[(scene_id, clip_id)] = np.load(path)
    return masks


def auc_from_curves(curves: Dict[int, List[np.ndarray]], clips, gts, masks=None) -> Tuple[float, Dict[int, float]]:
    """eval_COSKAD.py:226-253: per-transformation AUC on the concatenated clips, final AUC on the mean curve"""
    from sklearn.metrics import roc_auc_score
    gt_cat = []
    for (s, c, _f) in clips:
        g = gts[(s, c)]
        if masks is not None and (s, c) in masks:
            g = g[masks[(s, c)]]
        gt_cat.append(g)
    gt_cat = np.concatenate(gt_cat, axis=0)
    per_t, stack = {}, []
    for t, cl in curves.items():
        sc = np.concatenate(cl, axis=0)
        stack.append(sc)
        per_t[t] = float(roc_auc_score(gt_cat, sc))
    pds = np.mean(np.stack(stack, 0), 0)
    return float(roc_auc_score(gt_cat, pds)), per_t


class LitEncoder(LightningModule):
    """hyperbolic / Euclidean one-class encoder (flags ``hyperbolic``, ``static_center``)"""

    def __init__(self, args, hyperbolic: Optional[bool] = None) -> None:
        super().__init__()
        self.args = args
        self.hyperbolic = bool(args.hyperbolic if hyperbolic is None else hyperbolic)
        # tolerance of the Euclidean center init: 0.1 in models/hyperbolic_encoder.py:57, center_tolerance elsewhere
        self.eps = 0.1 if self.hyperbolic else float(getattr(args, 'center_tolerance', 1e-3))
        self.args.encoder_type = 'sts_gcn'
        self.model = STSE(input_dim=args.num_coords, layer_channels=list(getattr(args, 'channels', [32, 16, 32])),
                          hidden_dimension=args.h_dim, latent_dim=args.latent_dim, n_frames=args.dataset_seg_len,
                          n_joints=_joints(args), encoder_type='sts_gcn', projector=args.projector,
                          distance=getattr(args, 'distance', 'euclidean') if not self.hyperbolic else 'euclidean',
                          dropout=args.dropout)
        self.learning_rate = args.opt_lr
        self.batch_size = args.dataset_batch_size
        self.curvature = torch.tensor(-1.)
        self.temp: Optional[torch.Tensor] = None
        self._acc: Optional[torch.Tensor] = None
        self.centers: List[torch.Tensor] = []
        # distance 'mahalanobis' (models/euclidean_encoder_staticCenter.py:125-142,182-185; Euclidean encoder only): the
        # scatter matrix of the epoch's latents around the center is accumulated batch by batch (shard-additive float64
        # sums, all-reduced) instead of caching every latent batch (hidden_out_cache upstream) -- the center is constant
        # within an epoch, so the two are the same sum
        self.distance = 'euclidean' if self.hyperbolic else str(getattr(args, 'distance', 'euclidean')).lower()
        self._cov: Optional[torch.Tensor] = None

    # ---------------------------------------------------------------- forward (predict / validation)
    def forward(self, x):
        hidden_out = self.model(x[0])
        return hidden_out, x[1], x[2], x[3]

    @property
    def _flavour(self) -> int:
        return _lib.SCORE_POINCARE if self.hyperbolic else _lib.SCORE_EUCLID

    def _update_inv_cov(self) -> None:
        """compute_inv_cov_mat (euclidean_encoder_staticCenter.py:133-142) from the accumulated scatter matrix; in place, so a
        captured training step keeps reading the same buffer"""
        D = self.model.latent_dim
        cdist.allreduce_center_acc(self._cov)
        vi = gmath.inv_cov_finalize(self._cov, D)
        if self.model.inv_cov_matrix.shape == vi.shape and self.model.inv_cov_matrix.device == vi.device:
            self.model.inv_cov_matrix.copy_(vi)
        else:
            self.model.inv_cov_matrix = vi

    def _finalize_center(self, acc: torch.Tensor) -> torch.Tensor:
        cdist.allreduce_center_acc(acc)
        c = gmath.center_finalize(acc, self.model.latent_dim, self._flavour, eps=0.0 if self.hyperbolic else self.eps)
        if self.hyperbolic:
            assert bool((c < 1).all()), f'center is out of the ball\nc = {c}'     # hyperbolic_encoder.py:123
        return c

    # ---------------------------------------------------------------- center init (setup, :85-135)
    def setup(self, stage: Optional[str] = None) -> None:
        if stage != 'fit':
            return
        dev = self.trainer.device
        loader = self.trainer._data_connector._train_dataloader_source.dataloader()
        acc = gmath.center_accumulator(self.model.latent_dim, dev)
        self.model.eval()
        self.model.to(dev)
        with torch.no_grad():
            for batch in loader:
                data = (batch[0][0] if getattr(self.args, 'dataset_double_item', False) else batch[0]).to(dev)
                z, _ = self.model.encode_score(data)                       # fused eval kernel
                gmath.center_partial(gmath.expmap0_project(z) if self.hyperbolic else z, acc, self._flavour)
        c = self._finalize_center(acc)
        self.model.c = c
        self.temp = c
        self.centers.append(c.clone())
        if self.distance == 'mahalanobis':           # second pass: the scatter matrix needs the finished center (:125-126)
            self._cov = gmath.cov_accumulator(self.model.latent_dim, dev)
            with torch.no_grad():
                for batch in loader:
                    data = (batch[0][0] if getattr(self.args, 'dataset_double_item', False) else batch[0]).to(dev)
                    z, _ = self.model.encode_score(data)
                    gmath.cov_partial(z, c, self._cov)
            self._update_inv_cov()
        self.model.train()

    graph_safe = True

    def on_train_epoch_start(self) -> None:
        if self._acc is not None:
            self._acc.zero_()
        if self._cov is not None:                     # hidden_out_cache = [] upstream (:157-160)
            self._cov.zero_()

    # ---------------------------------------------------------------- training (:137-188)
    def training_step(self, batch, batch_idx):
        data = batch[0]
        hidden_out = self.model(data)
        loss_reg = calc_reg_loss(self.model)
        self.log('regularization', loss_reg)
        dynamic = not self.args.static_center
        if dynamic and self._acc is None:          # zeroed per epoch in on_train_epoch_start
            self._acc = gmath.center_accumulator(self.model.latent_dim, data.device)
        self.model.c = self.temp
        if self.hyperbolic:
            dist_c, hidden = gmath.poincare_score(hidden_out, self.model.c, True)   # fused expmap0 -> project -> dist(c, x)
            if dynamic:
                gmath.center_partial(hidden, self._acc, _lib.SCORE_POINCARE)
            loss_main = dist_c.mean()
            self.log('poincare_loss', loss_main)
            self.log('hyperlatent_norm', torch.linalg.norm(hidden, dim=-1).mean())
        else:
            if dynamic:
                gmath.center_partial(hidden_out.detach(), self._acc, _lib.SCORE_EUCLID)
            if self.distance == 'mahalanobis':       # :182-185
                gmath.cov_partial(hidden_out.detach(), self.model.c, self._cov)
                loss_main = gmath.mahalanobis(hidden_out, self.model.c, self.model.inv_cov_matrix)
            else:
                loss_main = F.mse_loss(hidden_out, self.model.c.expand_as(hidden_out))
            self.log('hypersphere_loss', loss_main)
        loss = loss_main + self.args.alpha * loss_reg
        self.log('loss', loss)
        return loss

    def training_epoch_end(self, outputs) -> None:
        if self.distance == 'mahalanobis' and self._cov is not None:      # on_train_epoch_end :145-148, before the center moves
            self._update_inv_cov()
        if self.args.static_center or self._acc is None:
            return
        c = self._finalize_center(self._acc)
        if self.hyperbolic:
            self.log('center/eucl', torch.norm(c, dim=-1))
            self.log('center/hyp', gmath.dist0(c.view(1, -1), k=-1.0)[0])
        # in place: a captured training step (Trainer(cuda_graph=True)) keeps reading this tensor
        if self.temp is not None and self.temp.shape == c.shape:
            self.temp.copy_(c)
        else:
            self.temp = c
        self.centers.append(c)

    def validation_step(self, batch, batch_idx):
        return self.forward(batch)

    def validation_epoch_end(self, outputs):
        hidden_out, trans, meta, frames = light_processing_data(outputs)
        return self.post_processing(hidden_out, trans, meta, frames)

    def configure_optimizers(self) -> Dict:
        optimizer = Adam(self.parameters(), lr=self.learning_rate, fused=True)          # no weight decay upstream (:199)
        if getattr(self.args, 'validation', False):
            sched = torch.optim.lr_scheduler.ReduceLROnPlateau(optimizer, mode='max', factor=0.2, patience=100, min_lr=1e-6)
            return {'optimizer': optimizer, 'lr_scheduler': sched, 'monitor': 'validation_auc'}
        sched = torch.optim.lr_scheduler.CosineAnnealingLR(optimizer, T_max=self.args.ae_epochs, eta_min=self.args.opt_lr)
        return {'optimizer': optimizer, 'lr_scheduler': sched}

    # ---------------------------------------------------------------- scoring (:219-305 / eval_COSKAD.py:140-253)
    def window_scores(self, hidden_out, validation: bool = False) -> torch.Tensor:
        """per-window anomaly score from raw latents.  Validation applies expmap0 WITHOUT project
        (hyperbolic_encoder.py:266); eval applies both (eval_COSKAD.py:195)."""
        dev = self.model.c.device if self.model.c.is_cuda else torch.device('cuda', torch.cuda.current_device())
        z = torch.as_tensor(hidden_out, dtype=torch.float32).to(dev)
        c = self.model.c.to(dev)
        if self.hyperbolic:
            x = gmath.expmap0(z, k=-1.0) if validation else gmath.expmap0_project(z)
            return gmath.dist(x, c, k=-1.0)
        if self.distance == 'mahalanobis':           # windows_based_loss_mahalanobis, utils/eval_utils.py:41-55
            return gmath.mahalanobis_score(z, c, self.model.inv_cov_matrix.to(dev))
        return gmath.euclid_score(z, c)

    def post_processing(self, hidden_out, trans, meta, frames, validation: bool = True):
        clips, gts = load_gt_table(self.args)
        nt = max(1, int(getattr(self.args, 'dataset_num_transform', 1)))
        scores = self.window_scores(hidden_out, validation=validation)
        curves = aggregate.score_and_aggregate(scores, trans, meta, frames, clips, nt,
                                               pad_size=-1 if validation else int(getattr(self.args, 'pad_size', -1)), gts=gts)
        auc, per_t = auc_from_curves(curves, clips, gts)
        self.log('validation_auc', auc)
        return auc


class LitAutoEncoder(LightningModule):
    """Euclidean auto-encoder (models/euclidean_autoencoder.py): lambda*mse(xhat,x) + mse(z,c) + alpha*reg"""

    def __init__(self, args) -> None:
        super().__init__()
        self.args = args
        self.eps = float(getattr(args, 'center_tolerance', 1e-3))
        self.model = STSAE(input_dim=args.num_coords, layer_channels=list(getattr(args, 'channels', [32, 16, 32])),
                           hidden_dimension=args.h_dim, latent_dim=args.latent_dim, n_frames=args.dataset_seg_len,
                           n_joints=_joints(args), encoder_type='sts_gcn', projector='linear', distance='euclidean',
                           dropout=args.dropout)
        self.learning_rate = args.opt_lr
        self.temp = None
        self._acc = None

    graph_safe = True          # static center, no per-epoch training state

    def forward(self, x):
        """(out, hidden, gt_data, trans, meta, frames): the 6-tuple consumed by light_processing_data"""
        z, xhat = self.model(x[0])
        return xhat, z, x[0], x[1], x[2], x[3]

    def setup(self, stage: Optional[str] = None) -> None:
        if stage != 'fit':
            return
        dev = self.trainer.device
        loader = self.trainer._data_connector._train_dataloader_source.dataloader()
        acc = gmath.center_accumulator(self.model.latent_dim, dev)
        self.model.eval().to(dev)
        with torch.no_grad():
            for batch in loader:
                z, _ = self.model.encode_score(batch[0].to(dev))
                gmath.center_partial(z, acc, _lib.SCORE_EUCLID)
        cdist.allreduce_center_acc(acc)
        self.model.c = gmath.center_finalize(acc, self.model.latent_dim, _lib.SCORE_EUCLID, eps=self.eps)
        self.temp = self.model.c
        self.model.train()

    def training_step(self, batch, batch_idx):
        data = batch[0]
        z, xhat = self.model(data)
        loss_rec = F.mse_loss(xhat, data)
        loss_hyp = F.mse_loss(z, self.model.c.expand_as(z))
        loss_reg = calc_reg_loss(self.model)
        loss = self.args.lambda_ * loss_rec + loss_hyp + self.args.alpha * loss_reg
        for k, v in (('loss', loss), ('reconstruction_loss', loss_rec), ('hypersphere_loss', loss_hyp), ('regularization', loss_reg)):
            self.log(k, v)
        return loss

    def validation_step(self, batch, batch_idx):
        return self.forward(batch)

    def validation_epoch_end(self, outputs):
        out, hidden_out, gt_data, trans, meta, frames = light_processing_data(outputs)
        return self.post_processing(out, hidden_out, gt_data, trans, meta, frames)

    def configure_optimizers(self) -> Dict:
        optimizer = Adam(self.parameters(), lr=self.learning_rate, fused=True)
        sched = torch.optim.lr_scheduler.CosineAnnealingLR(optimizer, T_max=self.args.ae_epochs, eta_min=self.args.opt_lr)
        return {'optimizer': optimizer, 'lr_scheduler': sched}

    def window_scores(self, out, hidden_out, gt_data, loss_type: str = 'rec', rec_loss_weight: float = 0.2) -> torch.Tensor:
        """utils/eval_utils.py:77-104: 'rec' (validation default), 'hyp' (eval_COSKAD.py:66-73), 'rec+hyp'"""
        dev = torch.device('cuda', torch.cuda.current_device())
        z = torch.as_tensor(hidden_out, dtype=torch.float32).to(dev)
        lat = gmath.euclid_score(z, self.model.c.to(dev))
        if loss_type == 'hyp':
            return lat
        o = torch.as_tensor(out, dtype=torch.float32).to(dev).reshape(z.shape[0], -1)
        g = torch.as_tensor(gt_data, dtype=torch.float32).to(dev).reshape(z.shape[0], -1)
        rec = gmath.euclid_score(o, g)
        return rec if loss_type == 'rec' else rec / rec_loss_weight + lat

    def post_processing(self, out, hidden_out, gt_data, trans, meta, frames, loss_type: str = 'rec'):
        clips, gts = load_gt_table(self.args)
        nt = max(1, int(getattr(self.args, 'dataset_num_transform', 1)))
        scores = self.window_scores(out, hidden_out, gt_data, loss_type)
        curves = aggregate.score_and_aggregate(scores, trans, meta, frames, clips, nt, gts=gts)
        auc, _ = auc_from_curves(curves, clips, gts)
        self.log('validation_auc', auc)
        return auc


def select_task(args):
    """the flag dispatch of train_COSKAD.py:36-55 / eval_COSKAD.py:60-91"""
    if args.use_decoder:
        return LitAutoEncoder
    if args.use_vae:
        from .spherical import LitSphericalVAE
        return LitSphericalVAE
    return LitEncoder
