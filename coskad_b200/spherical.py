"""Spherical VAE variant (flag ``use_vae``): models/sts/vae.py:13-169 (STSVAE) and models/spherical_vae.py
(LitEncoder), on the CUDA hot path.

  * encode: the fused eval kernel runs the encoder with a 9-row head (fc_mean stacked on fc_var); in training the
    per-layer kernels + two linear heads.  Z_mean = raw / ||raw|| (vae.py:81), kappa = softplus(raw_var) + 1 (vae.py:85).
  * reparameterisation: PowerSpherical(mu, kappa).rsample() = Householder_{e1->mu}([t, sqrt(1-t^2) v]) with
    t = 2 Beta(alpha, beta) - 1 and v uniform on S^{d-2}.  The noise (t, v) is drawn with torch's RNG (plumbing) and the
    transform runs in ``ps_sample_kernel`` at eval; in training the same transform is written with torch ops on the
    [B, 8] latents so that autograd provides the (implicit-reparameterisation) gradients -- a few hundred FLOP per
    window next to 8 MFLOP of network.
  * eval score: 1 - cos(mean_vector, Z) with Z a SAMPLE (spherical_vae.py:76, eval_COSKAD.py:81,191); a deterministic
    mode ``sample=False`` scores Z_mean instead.
power_spherical is un-vendored and unpinned upstream: the formulas follow its published source (parity unpinned).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.optim import Adam

from . import _lib, aggregate, dist as cdist, gmath, train
from .losses import calc_reg_loss
from .sts import STSAE, _check_x
from .trainer import LightningModule


class PowerSphericalQ(torch.distributions.Distribution):
    """the subset of power_spherical.PowerSpherical used by COSKAD: rsample(), entropy(), loc, scale"""
    arg_constraints = {}
    has_rsample = True

    def __init__(self, loc: torch.Tensor, scale: torch.Tensor, validate_args=None):
        self.loc, self.scale = loc, scale
        self.d = loc.shape[-1]
        super().__init__(batch_shape=loc.shape[:-1], event_shape=loc.shape[-1:], validate_args=False)

    def _alpha_beta(self):
        return (self.d - 1) / 2 + self.scale, torch.full_like(self.scale, (self.d - 1) / 2)

    def draw_noise(self):
        a, b = self._alpha_beta()
        t = 2 * torch.distributions.Beta(a, b).rsample() - 1              # reparameterised w.r.t. kappa
        g = torch.randn(self.loc.shape[:-1] + (self.d - 1,), device=self.loc.device, dtype=self.loc.dtype)
        return t, g / g.norm(dim=-1, keepdim=True)

    def rsample(self, sample_shape=torch.Size(), noise=None) -> torch.Tensor:
        if isinstance(sample_shape, (tuple, list)) and len(sample_shape) == 2 and torch.is_tensor(sample_shape[0]):
            noise, sample_shape = sample_shape, torch.Size()          # rsample((t, v)): explicit noise, positional
        if len(tuple(sample_shape)) != 0:
            raise NotImplementedError('COSKAD draws one sample per window (models/sts/vae.py:129)')
        t, v = self.draw_noise() if noise is None else noise
        if not (torch.is_grad_enabled() and (self.loc.requires_grad or t.requires_grad)):
            return ps_sample(self.loc, t, v)                               # CUDA kernel
        y = torch.cat((t.unsqueeze(-1), v * torch.sqrt(torch.clamp(1 - t.unsqueeze(-1) ** 2, 1e-7))), -1)
        u = torch.zeros_like(self.loc)
        u[..., 0] = 1.0
        u = u - self.loc
        u = u / (u.norm(dim=-1, keepdim=True) + 1e-5)
        return y - 2 * (y * u).sum(-1, keepdim=True) * u

    def entropy(self) -> torch.Tensor:
        a, b = self._alpha_beta()
        log_norm = -((a + b) * math.log(2) + torch.lgamma(a) - torch.lgamma(a + b) + b * math.log(math.pi))
        return -(log_norm + self.scale * (math.log(2) + torch.digamma(a) - torch.digamma(a + b)))


class HypersphericalUniformP(torch.distributions.Distribution):
    arg_constraints = {}

    def __init__(self, dim: int, device=None, dtype=None, validate_args=None):
        self.dim = dim
        super().__init__(batch_shape=torch.Size(), event_shape=torch.Size([dim + 1]), validate_args=False)

    def entropy(self) -> float:
        d = self.dim + 1
        return math.log(2) + (d / 2) * math.log(math.pi) - math.lgamma(d / 2)


@torch.distributions.kl.register_kl(PowerSphericalQ, HypersphericalUniformP)
def kl_divergence(q: PowerSphericalQ, p: HypersphericalUniformP) -> torch.Tensor:
    """KL(PS || U) = -H(PS) + H(U)   (power_spherical's registered KL, used at models/spherical_vae.py:92 through
    torch.distributions.kl.kl_divergence)"""
    return -q.entropy() + p.entropy()


def ps_sample(mu: torch.Tensor, t: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    mu2 = mu.detach().to(torch.float32).contiguous()
    B, d = mu2.shape
    t2 = t.detach().to(torch.float32).contiguous().view(-1)
    v2 = v.detach().to(torch.float32).contiguous()
    z = torch.empty_like(mu2)
    c = _lib.context(mu2.device.index if mu2.device.index is not None else torch.cuda.current_device())
    c.check(c.lib.coskad_ps_sample(c.h, mu2.data_ptr(), t2.data_ptr(), v2.data_ptr(), B, d, z.data_ptr(),
                                   _lib.stream_ptr(mu2.device)), 'coskad_ps_sample')
    return z


class STSVAE(STSAE):
    """models/sts/vae.py:13-169: distribution 'ps' (PowerSpherical, the configured one, config/UBnormal/spherical_vae.yaml:45)
    or 'normal' (diagonal Gaussian: fc_var has latent_dim outputs, Z_mean is not normalised, vae.py:79-85,107-108,160-169)"""

    def __init__(self, input_dim, layer_channels, hidden_dimension, latent_dim, n_frames, n_joints, encoder_type='sts_gcn',
                 projector='linear', distance='euclidean', dropout=0.0, bias=True, device='cpu', *,
                 projector_hidden_layers=None, distribution='ps'):
        self.distribution = distribution.lower()
        if self.distribution not in ('ps', 'normal'):
            raise ValueError(f'Distribution {self.distribution} not supported.')                      # vae.py:166
        if self.distribution == 'normal' and 2 * latent_dim > 16:
            raise NotImplementedError("distribution 'normal': the fused head holds 16 rows (fc_mean + fc_var = 2 * latent_dim), "
                                      f'latent_dim {latent_dim} needs {2 * latent_dim}')
        super().__init__(input_dim, layer_channels, hidden_dimension, latent_dim, n_frames, n_joints, encoder_type,
                         projector, distance, dropout, bias, device, projector_hidden_layers=projector_hidden_layers)
        if self.distribution == 'ps':
            self.mean_vector = None      # plain attribute for 'ps' upstream (spherical_vae.py:113), not a buffer

    def build_model(self) -> None:
        super().build_model()
        if self.distribution == 'normal':            # a buffer for 'normal', a plain attribute for 'ps' (vae.py:56-58)
            self.register_buffer('mean_vector', torch.zeros((1, self.latent_dim)))
        self.register_buffer('threshold_dist', torch.tensor(0, dtype=torch.float32))

    def _set_projector_type(self) -> None:
        if self.projector != 'linear':
            raise ValueError("projector 'mlp' is broken upstream (components.py:218) and not supported; use 'linear'")
        self.btlnk = nn.Identity()
        input_size = self.hidden_dimension * self.n_frames * self.n_joints
        self.fc_mean = nn.Linear(in_features=input_size, out_features=self.latent_dim)
        self.fc_var = nn.Linear(in_features=input_size, out_features=self.latent_dim if self.distribution == 'normal' else 1)

    def _head(self):
        w = torch.cat([self.fc_mean.weight, self.fc_var.weight], dim=0).detach().contiguous()
        b = torch.cat([self.fc_mean.bias, self.fc_var.bias], dim=0).detach().contiguous()
        self._head_keepalive = (w, b)
        return w, b, w.shape[0]                       # 'ps': latent_dim + 1 rows, 'normal': 2 * latent_dim

    def _sync_encoder(self, ctx) -> None:
        src = [self.fc_mean.weight, self.fc_mean.bias, self.fc_var.weight, self.fc_var.bias]
        tensors = list(self.encoder.parameters()) + list(self.encoder.buffers()) + src
        key = self._version_key(tensors)
        if key == self._enc_key:
            return
        hw, hb, rows = self._head()
        arr = self.encoder.layer_params_array()
        ctx.check(ctx.lib.coskad_set_encoder(ctx.h, len(self.encoder.model), arr, hw.data_ptr(), hb.data_ptr(), rows,
                                             _lib.stream_ptr(hw.device)), 'coskad_set_encoder')
        self._enc_key = key

    def encode(self, X: torch.Tensor, return_shape: bool = False):
        X = _check_x(X, self.input_dim, self.n_frames, self.n_joints)
        if self.training:
            H = train.encoder_features(self, X, True)
            raw = train.linear_reduce(H, self.fc_mean.weight, self.fc_mean.bias)
            Z_mean = raw / torch.norm(raw, dim=-1, keepdim=True) if self.distribution == 'ps' else raw   # vae.py:80-81
            Z_var = F.softplus(train.linear_reduce(H, self.fc_var.weight, self.fc_var.bias)) + 1   # vae.py:85
        else:
            raw9, _ = self.encode_score(X, _lib.SCORE_NONE)
            Z_mean = raw9[:, :self.latent_dim].contiguous()
            if self.distribution == 'ps':
                Z_mean = gmath.l2_normalize(Z_mean)
            Z_var = F.softplus(raw9[:, self.latent_dim:]) + 1
        if return_shape:
            return Z_mean, Z_var, torch.Size([X.shape[0], self.hidden_dimension, self.n_frames, self.n_joints, 1])
        return Z_mean, Z_var

    def reparameterize(self, Z_mean, Z_var):
        if self.distribution == 'normal':            # vae.py:107-109 (Z_var is used as the scale)
            return (torch.distributions.normal.Normal(Z_mean, Z_var),
                    torch.distributions.normal.Normal(torch.zeros_like(Z_mean), torch.ones_like(Z_var)))
        return PowerSphericalQ(loc=Z_mean, scale=torch.squeeze(Z_var, dim=-1)), HypersphericalUniformP(self.latent_dim - 1)

    @staticmethod
    def _rsample(q_Z, noise=None):
        """one reparameterised sample; ``noise``: the standard-normal draw ('normal') or the (t, v) pair ('ps')"""
        if isinstance(q_Z, torch.distributions.normal.Normal):
            return q_Z.rsample() if noise is None else q_Z.loc + q_Z.scale * noise
        return q_Z.rsample(noise=noise)

    def forward(self, X: torch.Tensor, noise=None):
        Z_mean, Z_var = self.encode(X)
        q_Z, p_Z = self.reparameterize(Z_mean, Z_var)
        Z = self._rsample(q_Z, noise)
        Xh = train.decode_forward(self, Z, self.training)
        return Z, Xh, (q_Z, p_Z, Z_var)

    @torch.no_grad()
    def cosine_scores(self, X: torch.Tensor, mean_vector: Optional[torch.Tensor] = None, sample: bool = False, noise=None):
        """eval score 1 - cos(mean_vector, Z); sample=False scores Z_mean with the fused kernel's cosine flavour"""
        mv = (self.mean_vector if mean_vector is None else mean_vector).view(-1)
        if not sample and self.distribution == 'ps':
            _, s = self.encode_score(X, _lib.SCORE_COSINE, center=mv, want_latent=False)
            return s
        if not sample:       # 'normal': the fused head holds 2 * latent_dim rows (mean | scale); score the mean rows
            return gmath.cosine_score(self.encode(X)[0], mv.to(X.device))
        Z_mean, Z_var = self.encode(X)
        Z = self._rsample(self.reparameterize(Z_mean, Z_var)[0], noise)
        return gmath.cosine_score(Z, mv.to(Z.device))


class STSVE(STSVAE):
    """old-kwarg constructor of the module the reference imports (models/spherical_vae.py:65-67)"""

    def __init__(self, c_in, h_dim, latent_dim, n_frames, dropout, n_joints, channels, distribution='ps',
                 decoder_channels=None, projector='linear', **kw):
        super().__init__(input_dim=c_in, layer_channels=list(channels), hidden_dimension=h_dim, latent_dim=latent_dim,
                         n_frames=n_frames, n_joints=n_joints, encoder_type='sts_gcn', projector=projector,
                         distance='euclidean', dropout=dropout, distribution=distribution)


class LitSphericalVAE(LightningModule):
    """models/spherical_vae.py:38-249"""

    def __init__(self, args) -> None:
        super().__init__()
        from .tasks import _joints
        self.args = args
        self.learning_rate = args.opt_lr
        self.phi, self.beta, self.gamma = args.phi, args.beta, args.gamma
        self.distribution = args.distribution
        self.warmup_counter = getattr(args, 'warmup_epochs', 0)
        self.updated_state_before_val = False
        self.model = STSVE(c_in=args.num_coords, h_dim=args.h_dim, latent_dim=args.latent_dim, n_frames=args.dataset_seg_len,
                           dropout=args.dropout, n_joints=_joints(args), channels=list(getattr(args, 'channels', [32, 16, 32])),
                           distribution=self.distribution, decoder_channels=getattr(args, 'decoder_channels', None),
                           projector=args.projector)
        self._acc = None

    def forward(self, x):
        hidden_out, rec, _ = self.model(x[0])
        return hidden_out, rec, x[1], x[2], x[3]

    def training_step(self, batch, batch_idx):
        data = batch[0]
        hidden_out, reconstructed_x, (q, p, z_var) = self.model(data)
        if self._acc is None:
            self._acc = gmath.center_accumulator(self.model.latent_dim, data.device)
        gmath.center_partial(hidden_out.detach(), self._acc, _lib.SCORE_COSINE)      # replaces latent_cache (:88)
        if self.distribution == 'normal':            # spherical_vae.py:89-92
            loss_kl = torch.distributions.kl.kl_divergence(q, p).sum(-1).mean()
        else:
            loss_kl = kl_divergence(q, p).mean()
        loss_rec = F.mse_loss(reconstructed_x, data)
        loss_reg = calc_reg_loss(self.model)
        loss_exp_dist = (1 / z_var).mean()
        loss = self.phi * loss_rec + self.args.alpha * loss_reg + self.beta * loss_kl + self.gamma * loss_exp_dist
        for k, v in (('loss', loss), ('reconstruction_loss', loss_rec), ('kl_loss', loss_kl), ('exp_dist_loss', loss_exp_dist),
                     ('regularization', loss_reg)):
            self.log(k, v)
        return loss

    def update_state(self) -> None:
        """mean_vector = empirical mean of the epoch's latent samples (spherical_vae.py:110-116), all-reduced"""
        if self._acc is None:
            return
        cdist.allreduce_center_acc(self._acc)
        mv = gmath.center_finalize(self._acc, self.model.latent_dim, _lib.SCORE_COSINE).view(1, -1)
        if self.distribution == 'normal' and torch.is_tensor(self.model.mean_vector) and self.model.mean_vector.device == mv.device:
            self.model.mean_vector.copy_(mv)          # a registered buffer for 'normal' (vae.py:57): stays in the state_dict
        else:
            self.model.mean_vector = mv
        if self.warmup_counter > 0:
            self.warmup_counter -= 1
        self._acc = None

    def on_validation_start(self) -> None:
        self.update_state()
        self.updated_state_before_val = True

    def on_train_epoch_end(self) -> None:
        if not self.updated_state_before_val:
            self.update_state()
        self.updated_state_before_val = False

    def validation_step(self, batch, batch_idx):
        return self.forward(batch)

    def validation_epoch_end(self, outputs):
        from .tasks import light_processing_data
        hidden_out, trans, meta, frames = light_processing_data(outputs)
        return self.post_processing(hidden_out, trans, meta, frames)

    def configure_optimizers(self) -> Dict:
        optimizer = Adam(self.parameters(), lr=self.learning_rate, fused=True)
        if getattr(self.args, 'validation', False):
            sched = torch.optim.lr_scheduler.ReduceLROnPlateau(optimizer, mode='max', factor=0.2, patience=2, min_lr=1e-6)
            return {'optimizer': optimizer, 'lr_scheduler': sched, 'monitor': 'validation_auc'}
        sched = torch.optim.lr_scheduler.CosineAnnealingLR(optimizer, T_max=self.args.ae_epochs, eta_min=self.args.opt_lr)
        return {'optimizer': optimizer, 'lr_scheduler': sched}

    def window_scores(self, hidden_out, validation: bool = False) -> torch.Tensor:
        dev = torch.device('cuda', torch.cuda.current_device())
        z = torch.as_tensor(hidden_out, dtype=torch.float32).to(dev)
        return gmath.cosine_score(z, self.model.mean_vector.view(-1).to(dev))       # eval_COSKAD.py:81,191

    def post_processing(self, hidden_out, trans, meta, frames):
        from .tasks import auc_from_curves, load_gt_table
        clips, gts = load_gt_table(self.args)
        nt = max(1, int(getattr(self.args, 'dataset_num_transform', 1)))
        curves = aggregate.score_and_aggregate(self.window_scores(hidden_out), trans, meta, frames, clips, nt, gts=gts)
        auc, _ = auc_from_curves(curves, clips, gts)
        self.log('validation_auc', auc)
        return auc
