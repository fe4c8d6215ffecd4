"""Oracle (TEST INFRASTRUCTURE): torch-CPU restatement of COSKAD's STS-GCN networks.

Functional, state-dict driven; it issues the same ATen ops as the reference modules
(einsum, conv2d 1x1, batch_norm, prelu, linear) so on CPU it agrees with them to the last
bit or two, and it is what ``bench.py`` times as the ``cpu_baseline`` ("port").

State-dict keys are the ones the in-tree reference classes produce
(``encoder.model.{i}.gcn.A`` ... ``btlnk.weight`` ... ``decoder.model.{i}...``), so a dict made
here loads into ``models.sts.ae.STSE/STSAE`` unchanged (that is how ``gen_golden.py`` pins it).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
BN_EPS = 1e-5        # nn.BatchNorm2d default  (models/graph_layers/stsgcn.py:65,77)
BN_MOMENTUM = 0.1


# --------------------------------------------------------------------------- layer
def graph_contract(x: Tensor, A: Tensor, T: Tensor) -> Tensor:
    """ConvTemporalGraphical.forward -- models/graph_layers/stsgcn.py:143-156.

    x [N,C,T,V]; T [V,T,T] mixes frames per joint, then A [T,V,V] mixes joints per frame."""
    x = torch.einsum('nctv,vtq->ncqv', x, T).contiguous()      # stsgcn.py:154
    x = torch.einsum('nctv,tvw->nctw', x, A).contiguous()      # stsgcn.py:155
    return x


def _conv_bn(x: Tensor, sd: Dict[str, Tensor], pfx: str, training: bool,
             new_stats: Optional[Dict[str, Tensor]]) -> Tensor:
    """1x1 Conv2d followed by BatchNorm2d -- stsgcn.py:56-66 (tcn) / :71-77 (residual)."""
    y = F.conv2d(x, sd[pfx + '.0.weight'], sd.get(pfx + '.0.bias'))
    rm, rv = sd[pfx + '.1.running_mean'], sd[pfx + '.1.running_var']
    if training:
        rm, rv = rm.clone(), rv.clone()
    y = F.batch_norm(y, rm, rv, sd[pfx + '.1.weight'], sd[pfx + '.1.bias'],
                     training=training, momentum=BN_MOMENTUM, eps=BN_EPS)
    if training and new_stats is not None:
        new_stats[pfx + '.1.running_mean'] = rm
        new_stats[pfx + '.1.running_var'] = rv
    return y


def st_gcnn_layer(x: Tensor, sd: Dict[str, Tensor], pfx: str, training: bool = False,
                  new_stats: Optional[Dict[str, Tensor]] = None, prelu_mask: Optional[Tensor] = None) -> Tensor:
    """ST_GCNN_layer.forward -- stsgcn.py:94-116 (dropout p=0 in every config; emb branch unused).

    prelu_mask (test aid): use this boolean "positive branch" pattern instead of ``pre > 0``.  fp32
    implementations legitimately disagree on the sign of pre-activations that are within rounding of
    zero; fixing the pattern lets the gradient kernels be compared exactly."""
    if (pfx + '.residual.0.weight') in sd:
        res = _conv_bn(x, sd, pfx + '.residual', training, new_stats)       # stsgcn.py:106
    else:
        res = x                                                               # nn.Identity, :80
    g = graph_contract(x, sd[pfx + '.gcn.A'], sd[pfx + '.gcn.T'])            # :107
    y = _conv_bn(g, sd, pfx + '.tcn', training, new_stats)                   # :108
    pre = y + res
    if prelu_mask is not None:
        return torch.where(prelu_mask, pre, sd[pfx + '.prelu.weight'] * pre)
    return F.prelu(pre, sd[pfx + '.prelu.weight'])                           # :109-110


def layer_stack(x: Tensor, sd: Dict[str, Tensor], pfx: str, training: bool = False,
                new_stats: Optional[Dict[str, Tensor]] = None,
                return_all: bool = False, prelu_masks: Optional[List[Tensor]] = None):
    """Encoder.forward / Decoder.forward -- models/common/components.py:94-105, 168-179."""
    acts = []
    i = 0
    while f'{pfx}.model.{i}.gcn.A' in sd:
        x = st_gcnn_layer(x, sd, f'{pfx}.model.{i}', training, new_stats,
                          None if prelu_masks is None else prelu_masks[i])
        acts.append(x)
        i += 1
    return (x, acts) if return_all else x


# --------------------------------------------------------------------------- networks
def stse_forward(x: Tensor, sd: Dict[str, Tensor], training: bool = False,
                 new_stats: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """STSE.encode/forward -- models/sts/ae.py:76-121.

    The unsqueeze/permute/view round trip at ae.py:89-93 is the identity for M=1; the flatten at
    :96-100 is (c,t,v) order, i.e. ``h.reshape(B,-1)`` of the contiguous [B,C,T,V] activation."""
    assert x.dim() == 4, 'Input tensor must have shape [batch_size, input_dim, n_frames, n_joints]'
    h = layer_stack(x, sd, 'encoder', training, new_stats)
    h = h.reshape(h.shape[0], -1)
    return F.linear(h, sd['btlnk.weight'], sd.get('btlnk.bias'))


def stsae_forward(x: Tensor, sd: Dict[str, Tensor], training: bool = False,
                  new_stats: Optional[Dict[str, Tensor]] = None) -> Tuple[Tensor, Tensor]:
    """STSAE.forward -- models/sts/ae.py:233-250; returns (Z, X_hat) like the in-tree class."""
    z = stse_forward(x, sd, training, new_stats)
    return z, stsae_decode(z, sd, x.shape, training, new_stats)


def stsae_decode(z: Tensor, sd: Dict[str, Tensor], x_shape, training: bool = False,
                 new_stats: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """STSAE.decode -- models/sts/ae.py:210-230 (rev_btlnk, view [B,H,T,V], Decoder)."""
    B, _, Tn, Vn = x_shape
    h = F.linear(z, sd['rev_btlnk.weight'], sd['rev_btlnk.bias'])
    h = h.view(B, -1, Tn, Vn)
    return layer_stack(h, sd, 'decoder', training, new_stats)


def stsvae_encode(x: Tensor, sd: Dict[str, Tensor], training: bool = False,
                  new_stats: Optional[Dict[str, Tensor]] = None,
                  distribution: str = 'ps') -> Tuple[Tensor, Tensor]:
    """STSVAE.encode -- models/sts/vae.py:63-91 (btlnk = Identity for projector 'linear', :151)."""
    h = layer_stack(x, sd, 'encoder', training, new_stats)
    h = h.reshape(h.shape[0], -1)
    z_mean = F.linear(h, sd['fc_mean.weight'], sd['fc_mean.bias'])
    if distribution == 'ps':
        z_mean = z_mean / torch.norm(z_mean, dim=-1, keepdim=True)           # vae.py:81
    z_var = F.softplus(F.linear(h, sd['fc_var.weight'], sd['fc_var.bias'])) + 1   # vae.py:85
    return z_mean, z_var


# --------------------------------------------------------------------------- BN folding (eval)
def fold_layer_eval(sd: Dict[str, Tensor], pfx: str) -> Tuple[Tensor, Tensor, Tensor]:
    """Eval-mode algebra the CUDA path relies on: BN1(W1 g + b1) + BN2(W2 x + b2)
    = W1' g + W2' x + b'.  Returns (W1' [Co,Ci], W2' [Co,Ci], b' [Co]) in float64-rounded-to-f32."""
    def fold(p):
        w = sd[p + '.0.weight'].double().flatten(1)
        b = sd[p + '.0.bias'].double() if (p + '.0.bias') in sd else torch.zeros(w.shape[0], dtype=torch.float64)
        s = sd[p + '.1.weight'].double() / torch.sqrt(sd[p + '.1.running_var'].double() + BN_EPS)
        return w * s[:, None], (b - sd[p + '.1.running_mean'].double()) * s + sd[p + '.1.bias'].double()
    w1, c1 = fold(pfx + '.tcn')
    w2, c2 = fold(pfx + '.residual')
    return w1.float(), w2.float(), (c1 + c2).float()


# --------------------------------------------------------------------------- parameters
def _conv_init(co: int, ci: int, g: torch.Generator) -> Tuple[Tensor, Tensor]:
    bound = 1.0 / math.sqrt(ci)          # kaiming_uniform(a=sqrt(5)) on a 1x1 kernel == U(+-1/sqrt(fan_in))
    w = (torch.rand(co, ci, 1, 1, generator=g) * 2 - 1) * bound
    b = (torch.rand(co, generator=g) * 2 - 1) * bound
    return w, b


def _layer_params(sd: Dict[str, Tensor], pfx: str, ci: int, co: int, Tn: int, Vn: int,
                  g: torch.Generator, randomize_bn: bool) -> None:
    sd[pfx + '.gcn.A'] = (torch.rand(Tn, Vn, Vn, generator=g) * 2 - 1) / math.sqrt(Vn)   # stsgcn.py:134-136
    sd[pfx + '.gcn.T'] = (torch.rand(Vn, Tn, Tn, generator=g) * 2 - 1) / math.sqrt(Tn)   # stsgcn.py:138-140
    for br in ('tcn', 'residual'):
        w, b = _conv_init(co, ci, g)
        sd[f'{pfx}.{br}.0.weight'], sd[f'{pfx}.{br}.0.bias'] = w, b
        if randomize_bn:      # make the fold non-trivial
            sd[f'{pfx}.{br}.1.weight'] = torch.rand(co, generator=g) + 0.5
            sd[f'{pfx}.{br}.1.bias'] = torch.randn(co, generator=g) * 0.1
            sd[f'{pfx}.{br}.1.running_mean'] = torch.randn(co, generator=g) * 0.2
            sd[f'{pfx}.{br}.1.running_var'] = torch.rand(co, generator=g) + 0.5
        else:
            sd[f'{pfx}.{br}.1.weight'] = torch.ones(co)
            sd[f'{pfx}.{br}.1.bias'] = torch.zeros(co)
            sd[f'{pfx}.{br}.1.running_mean'] = torch.zeros(co)
            sd[f'{pfx}.{br}.1.running_var'] = torch.ones(co)
        sd[f'{pfx}.{br}.1.num_batches_tracked'] = torch.zeros((), dtype=torch.long)
    sd[pfx + '.prelu.weight'] = torch.full((1,), 0.25)                                  # nn.PReLU default


def init_state_dict(kind: str = 'stse', input_dim: int = 2, layer_channels: Sequence[int] = (32, 16, 32),
                    hidden_dimension: int = 64, latent_dim: int = 16, n_frames: int = 12, n_joints: int = 17,
                    seed: int = 0, randomize_bn: bool = True, distribution: str = 'ps') -> Dict[str, Tensor]:
    """Seeded random parameters with the reference's init distributions and key names
    (kind: 'stse' | 'stsae' | 'stsvae'; distribution 'normal': fc_var has latent_dim rows and mean_vector is a buffer,
    models/sts/vae.py:56-58,160-169).  Deterministic for a given torch build (CPU generator)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {'c': torch.zeros(latent_dim)}
    chans = [input_dim] + list(layer_channels) + [hidden_dimension]
    for i in range(len(chans) - 1):
        _layer_params(sd, f'encoder.model.{i}', chans[i], chans[i + 1], n_frames, n_joints, g, randomize_bn)
    F_ = hidden_dimension * n_frames * n_joints
    bound = 1.0 / math.sqrt(F_)
    if kind in ('stse', 'stsae'):
        sd['btlnk.weight'] = (torch.rand(latent_dim, F_, generator=g) * 2 - 1) * bound
        sd['btlnk.bias'] = (torch.rand(latent_dim, generator=g) * 2 - 1) * bound
    if kind == 'stsvae':
        sd['fc_mean.weight'] = (torch.rand(latent_dim, F_, generator=g) * 2 - 1) * bound
        sd['fc_mean.bias'] = (torch.rand(latent_dim, generator=g) * 2 - 1) * bound
        vr = latent_dim if distribution == 'normal' else 1
        sd['fc_var.weight'] = (torch.rand(vr, F_, generator=g) * 2 - 1) * bound
        sd['fc_var.bias'] = (torch.rand(vr, generator=g) * 2 - 1) * bound
        sd['threshold_dist'] = torch.zeros(())
        if distribution == 'normal':
            sd['mean_vector'] = torch.zeros(1, latent_dim)
    if kind in ('stsae', 'stsvae'):
        b2 = 1.0 / math.sqrt(latent_dim)
        sd['rev_btlnk.weight'] = (torch.rand(F_, latent_dim, generator=g) * 2 - 1) * b2
        sd['rev_btlnk.bias'] = (torch.rand(F_, generator=g) * 2 - 1) * b2
        dch = [hidden_dimension] + list(layer_channels)[::-1] + [input_dim]
        for i in range(len(dch) - 1):
            _layer_params(sd, f'decoder.model.{i}', dch[i], dch[i + 1], n_frames, n_joints, g, randomize_bn)
    return sd


def synth_windows(n: int, seed: int = 999, shape: str = 'ubnormal', n_coords: int = 2,
                  n_frames: int = 12, n_joints: int = 17) -> Tensor:
    """Synthetic pose windows.  'ubnormal': robust-scaled coords ~N(0,0.4^2) clamped to
    +-3 with 2% of (t,v) joints zeroed (utils/data.py:374-383 zeros = missing joints); 'stc':
    per-window centre U(-1,1) + sigma 0.1, clamped to [-1,1] (utils/dataset_utils.py:36-42)."""
    g = torch.Generator().manual_seed(seed)
    if shape == 'ubnormal':
        x = (torch.randn(n, n_coords, n_frames, n_joints, generator=g) * 0.4).clamp_(-3, 3)
        drop = torch.rand(n, 1, n_frames, n_joints, generator=g) < 0.02
        x = x.masked_fill(drop, 0.0)
    elif shape == 'stc':
        ctr = torch.rand(n, n_coords, 1, 1, generator=g) * 2 - 1
        x = (ctr + 0.1 * torch.randn(n, n_coords, n_frames, n_joints, generator=g)).clamp_(-1, 1)
    else:
        raise ValueError(shape)
    return x.contiguous()


def calc_reg_loss(sd_params: List[Tuple[str, Tensor]]) -> Tensor:
    """utils/model_utils.py:90-103 -- 0.5*sum ||p||^2 over non-'bias' tensors / number of tensors."""
    ps = [p for n, p in sd_params if 'bias' not in n]
    reg = None
    for p in ps:
        reg = 0.5 * torch.sum(p ** 2) if reg is None else reg + 0.5 * p.norm(2) ** 2
    return reg / len(ps)
