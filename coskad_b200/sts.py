"""Host-side mirror of the reference's network classes, backed by libcoskad_b200.so.

Same class names, constructor arguments, parameter / buffer names (so ``state_dict``s interchange
with the reference: ``encoder.model.{i}.gcn.A`` ... ``btlnk.weight`` ... ``c``) and return values as
    models/sts/ae.py:12   STSE     models/sts/ae.py:168  STSAE     models/sts/vae.py:13  STSVAE
    models/common/components.py:45 Encoder, :109 Decoder
    models/graph_layers/stsgcn.py:9 ST_GCNN_layer, :120 ConvTemporalGraphical
The torch ``nn`` sub-modules (Conv2d, BatchNorm2d, PReLU, Linear) are used as PARAMETER CONTAINERS
only -- their ``forward`` is never called.  All arithmetic runs in the hand-written sm_100a kernels
through the C ABI; without the library or a B200 the calls raise (no fallback).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List, Optional, Tuple, Union

import torch
import torch.nn as nn

from . import _lib
from ._lib import LayerParams


# ------------------------------------------------------------------------------------------------
# parameter containers (reference module tree)
class ConvTemporalGraphical(nn.Module):
    """models/graph_layers/stsgcn.py:120-140 (parameters and init only)."""

    def __init__(self, time_dim: int, joints_dim: int) -> None:
        super().__init__()
        self.A = nn.Parameter(torch.empty(time_dim, joints_dim, joints_dim))
        stdv = 1. / math.sqrt(self.A.size(1))
        self.A.data.uniform_(-stdv, stdv)
        self.T = nn.Parameter(torch.empty(joints_dim, time_dim, time_dim))
        stdv = 1. / math.sqrt(self.T.size(1))
        self.T.data.uniform_(-stdv, stdv)


class ST_GCNN_layer(nn.Module):
    """models/graph_layers/stsgcn.py:9-91 (module tree only; the arithmetic is in the CUDA kernels)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size, stride: int, time_dim: int,
                 joints_dim: int, dropout: float, bias: bool = True, emb_dim: Optional[int] = None) -> None:
        super().__init__()
        assert kernel_size[0] % 2 == 1 and kernel_size[1] % 2 == 1
        if tuple(kernel_size) != (1, 1) or stride != 1:
            raise NotImplementedError('coskad_b200 implements the (1,1)/stride-1 layer every COSKAD config uses')
        if emb_dim is not None:
            raise NotImplementedError('emb_dim branch (stsgcn.py:84-91) is unused by COSKAD and not implemented')
        if dropout != 0:
            raise NotImplementedError('dropout > 0: every reference config uses dropout 0 '
                                      '(config/*/*.yaml); the CUDA training kernels implement p = 0 only')
        self.in_channels, self.out_channels = in_channels, out_channels
        self.time_dim, self.joints_dim = time_dim, joints_dim
        self.gcn = ConvTemporalGraphical(time_dim, joints_dim)
        self.tcn = nn.Sequential(nn.Conv2d(in_channels, out_channels, (1, 1), (1, 1), (0, 0), bias=bool(bias)),
                                 nn.BatchNorm2d(out_channels), nn.Dropout(dropout, inplace=True))
        if in_channels != out_channels:
            self.residual = nn.Sequential(nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=(1, 1), bias=bool(bias)),
                                          nn.BatchNorm2d(out_channels))
        else:
            self.residual = nn.Identity()
        self.prelu = nn.PReLU()

    def layer_params(self) -> LayerParams:
        """device pointers of this layer's tensors as a coskad_layer_params"""
        p = LayerParams()
        p.c_in, p.c_out = self.in_channels, self.out_channels
        conv, bn = self.tcn[0], self.tcn[1]
        p.A, p.T = self.gcn.A.data_ptr(), self.gcn.T.data_ptr()
        p.w1 = conv.weight.data_ptr()
        p.b1 = conv.bias.data_ptr() if conv.bias is not None else None
        p.bn1_w, p.bn1_b = bn.weight.data_ptr(), bn.bias.data_ptr()
        p.bn1_rm, p.bn1_rv = bn.running_mean.data_ptr(), bn.running_var.data_ptr()
        if isinstance(self.residual, nn.Identity):
            p.w2 = None
        else:
            conv2, bn2 = self.residual[0], self.residual[1]
            p.w2 = conv2.weight.data_ptr()
            p.b2 = conv2.bias.data_ptr() if conv2.bias is not None else None
            p.bn2_w, p.bn2_b = bn2.weight.data_ptr(), bn2.bias.data_ptr()
            p.bn2_rm, p.bn2_rv = bn2.running_mean.data_ptr(), bn2.running_var.data_ptr()
        p.prelu = self.prelu.weight.data_ptr()
        return p


class _LayerStack(nn.Module):
    def _build(self, chans: List[int], n_frames: int, n_joints: int, dropout: float, bias: bool) -> None:
        layers = [ST_GCNN_layer(ci, co, (1, 1), 1, n_frames, n_joints, dropout, bias)
                  for ci, co in zip(chans[:-1], chans[1:])]
        self.model = nn.Sequential(*layers)

    def layer_params_array(self):
        arr = (LayerParams * len(self.model))()
        for i, l in enumerate(self.model):
            arr[i] = l.layer_params()
        return arr

    def forward(self, X):  # pragma: no cover - the stacks are driven by STSE/STSAE
        raise RuntimeError('Encoder/Decoder are parameter containers: call the owning STSE/STSAE/STSVAE')


class Encoder(_LayerStack):
    """models/common/components.py:45-105"""

    def __init__(self, input_dim, layer_channels, hidden_dimension, n_frames, n_joints, dropout, bias=True, device='cpu'):
        super().__init__()
        self._build([input_dim] + list(layer_channels) + [hidden_dimension], n_frames, n_joints, dropout, bias)


class Decoder(_LayerStack):
    """models/common/components.py:109-179"""

    def __init__(self, output_dim, layer_channels, hidden_dimension, n_frames, n_joints, dropout, bias=True, device='cpu'):
        super().__init__()
        self._build([hidden_dimension] + list(layer_channels)[::-1] + [output_dim], n_frames, n_joints, dropout, bias)


# ------------------------------------------------------------------------------------------------
def _check_x(X: torch.Tensor, n_coords: int, n_frames: int, n_joints: int) -> torch.Tensor:
    assert len(X.shape) == 4, f'Input tensor must have shape [batch_size, input_dim, n_frames, n_joints]. Got {X.shape}'
    if not X.is_cuda:
        raise _lib.CoskadError('coskad_b200 runs on a B200 only: the input tensor is not on a CUDA device (no CPU fallback)')
    if tuple(X.shape[1:]) != (n_coords, n_frames, n_joints):
        raise ValueError(f'expected [B,{n_coords},{n_frames},{n_joints}], got {tuple(X.shape)}')
    if X.dtype != torch.float32:
        X = X.float()
    return X.contiguous()


class STSE(nn.Module):
    """models/sts/ae.py:12-165 -- STS-GCN encoder + linear bottleneck + center buffer ``c``."""

    def __init__(self, input_dim: int, layer_channels: List[int], hidden_dimension: int, latent_dim: int,
                 n_frames: int, n_joints: int, encoder_type: str = 'sts_gcn', projector: str = 'linear',
                 distance: str = 'euclidean', dropout: float = 0., bias: bool = True,
                 device: Union[str, torch.device] = 'cpu', *, projector_hidden_layers: Optional[List[int]] = None) -> None:
        super().__init__()
        self.input_dim, self.layer_channels = input_dim, list(layer_channels)
        self.hidden_dimension, self.latent_dim = hidden_dimension, latent_dim
        self.n_frames, self.n_joints = n_frames, n_joints
        self.encoder_type, self.projector = encoder_type.lower(), projector.lower()
        self.projector_hidden_layers = projector_hidden_layers
        self.distance, self.dropout, self.bias, self.device = distance.lower(), dropout, bias, device
        self._ctx: Optional[_lib.Context] = None
        self.fused_impl = 1      # 1: tcgen05 channel mixing (default), 0: all-FP32 CUDA-core kernel
        self._enc_key = None
        self._dec_key = None
        self.build_model()

    # -- construction (ae.py:60-72,124-164)
    def build_model(self) -> None:
        self._set_encoder_type()
        self._set_projector_type()
        self.register_buffer('c', torch.zeros(self.latent_dim))
        if self.distance == 'mahalanobis':
            self.register_buffer('inv_cov_matrix', torch.zeros((self.latent_dim, self.latent_dim)))

    def _set_encoder_type(self) -> None:
        if self.encoder_type != 'sts_gcn':
            raise ValueError(f'Encoder type {self.encoder_type} not supported by coskad_b200 '
                             '(only the sts_gcn hot path is implemented; the ablation encoders are out of scope).')
        self.encoder = Encoder(self.input_dim, self.layer_channels, self.hidden_dimension, self.n_frames,
                               self.n_joints, self.dropout, self.bias, self.device)

    def _set_projector_type(self) -> None:
        input_size = self.hidden_dimension * self.n_frames * self.n_joints
        if self.projector == 'linear':
            self.btlnk = nn.Linear(in_features=input_size, out_features=self.latent_dim, bias=bool(self.bias))
        elif self.projector == 'mlp':
            # upstream MLP.build_model raises UnboundLocalError (models/common/components.py:218)
            raise ValueError("projector 'mlp' is broken upstream (components.py:218) and not supported; use 'linear'")
        else:
            raise ValueError(f'Projector type {self.projector} not supported.')

    def train(self, mode: bool = True):
        """every train/eval switch drops the folded-weight cache: a CUDA-graph replay of the training step updates the
        parameters and the BatchNorm statistics on the device without bumping any ``tensor._version``, so the version key
        of ``_sync_encoder`` alone would let an eval pass score with the weights of the previous fold"""
        self._enc_key = self._dec_key = None
        return super().train(mode)

    # -- C-ABI plumbing
    def _context(self, X: torch.Tensor) -> _lib.Context:
        dev = X.device.index if X.device.index is not None else torch.cuda.current_device()
        if self._ctx is None or self._ctx.device != dev:
            self._ctx = _lib.Context(dev, self.n_frames, self.n_joints)
            self._enc_key = self._dec_key = None
        self._ctx.check(self._ctx.lib.coskad_set_fused_impl(self._ctx.h, int(self.fused_impl)), 'coskad_set_fused_impl')
        return self._ctx

    @staticmethod
    def _version_key(tensors) -> Tuple:
        return tuple((t.data_ptr(), t._version) for t in tensors)

    def _head(self) -> Tuple[torch.Tensor, Optional[torch.Tensor], int]:
        return self.btlnk.weight, self.btlnk.bias, self.latent_dim

    def _sync_encoder(self, ctx: _lib.Context) -> None:
        hw, hb, rows = self._head()
        tensors = list(self.encoder.parameters()) + list(self.encoder.buffers()) + [hw] + ([hb] if hb is not None else [])
        key = self._version_key(tensors)
        if key == self._enc_key:
            return
        for t in tensors:
            if not t.is_cuda or not t.is_contiguous() or (t.is_floating_point() and t.dtype != torch.float32):
                raise _lib.CoskadError('model parameters must be contiguous float32 CUDA tensors (call .cuda())')
        arr = self.encoder.layer_params_array()
        rc = ctx.lib.coskad_set_encoder(ctx.h, len(self.encoder.model), arr, hw.data_ptr(),
                                        hb.data_ptr() if hb is not None else None, rows, _lib.stream_ptr(hw.device))
        ctx.check(rc, 'coskad_set_encoder')
        self._enc_key = key

    # -- forward (ae.py:76-121)
    def encode(self, X: torch.Tensor, return_shape: bool = False):
        X = _check_x(X, self.input_dim, self.n_frames, self.n_joints)
        if self.training:
            from .train import stse_train_forward
            Z = stse_train_forward(self, X)
        else:
            Z, _ = self.encode_score(X, _lib.SCORE_NONE)
        if return_shape:
            B = X.shape[0]
            return Z, torch.Size([B, self.hidden_dimension, self.n_frames, self.n_joints, 1])
        return Z

    def forward(self, X: torch.Tensor) -> torch.Tensor:
        return self.encode(X)

    @torch.no_grad()
    def encode_score(self, X: torch.Tensor, flavour: int = _lib.SCORE_NONE, center: Optional[torch.Tensor] = None,
                     want_latent: bool = True, score_out: Optional[torch.Tensor] = None
                     ) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
        """Fused eval hot path: one kernel from pose windows to (raw latent, anomaly score).

        ``center`` defaults to the buffer ``c``.  Returns (Z [B, head_rows] or None, score [B] or None)."""
        X = _check_x(X, self.input_dim, self.n_frames, self.n_joints)
        ctx = self._context(X)
        self._sync_encoder(ctx)
        B = X.shape[0]
        rows = self._head()[2]
        Z = torch.empty((B, rows), device=X.device, dtype=torch.float32) if want_latent else None
        score = None
        cen = None
        if flavour != _lib.SCORE_NONE:
            cen = (self.c if center is None else center).to(device=X.device, dtype=torch.float32).contiguous().view(-1)
            if score_out is not None:
                assert score_out.is_cuda and score_out.dtype == torch.float32 and score_out.is_contiguous() \
                    and score_out.numel() == B
                score = score_out
            else:
                score = torch.empty((B,), device=X.device, dtype=torch.float32)
        rc = ctx.lib.coskad_encode_score_fwd(ctx.h, int(flavour), X.data_ptr(), _lib._ptr(cen), B, _lib._ptr(Z),
                                             _lib._ptr(score), _lib.stream_ptr(X.device))
        ctx.check(rc, 'coskad_encode_score_fwd')
        return Z, score


    @torch.no_grad()
    def encode_score_traj(self, traj: torch.Tensor, win_row: torch.Tensor, trans: Optional[torch.Tensor] = None,
                          mats: Optional[torch.Tensor] = None, flavour: int = _lib.SCORE_NONE,
                          center: Optional[torch.Tensor] = None, want_latent: bool = True,
                          score_out: Optional[torch.Tensor] = None
                          ) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
        """Fused eval hot path fed from trajectories: window construction (utils/preprocessing.py:58-89) and the
        test-time affine transforms (utils/dataset.py:65-74, utils/dataset_utils.py:255-310) happen inside the kernel.

        ``traj`` [rows, 2*V] float32 (x0,y0,x1,y1,.. per frame: the reference's scaled ``trajectory.coordinates`` rows,
        persons concatenated); window i = the ``n_frames`` consecutive rows from ``win_row[i]``; ``trans`` [N] indexes
        ``mats`` [n_mats, 2, 3] (rows 0,1 of the affine matrices).  Returns (Z, score) like :meth:`encode_score`."""
        if traj.dim() != 2 or traj.shape[1] != 2 * self.n_joints or not traj.is_cuda:
            raise ValueError(f'traj must be a CUDA tensor [rows, {2 * self.n_joints}], got {tuple(traj.shape)} on {traj.device}')
        traj = traj.to(torch.float32).contiguous()
        win_row = win_row.to(device=traj.device, dtype=torch.int64).contiguous()
        N = win_row.numel()
        if (trans is None) != (mats is None):
            raise ValueError('trans and mats must be given together')
        n_mats = 0
        if trans is not None:
            trans = trans.to(device=traj.device, dtype=torch.int32).contiguous()
            mats = mats.to(device=traj.device, dtype=torch.float32).reshape(-1, 6).contiguous()
            n_mats = mats.shape[0]
            if trans.numel() != N:
                raise ValueError('trans must have one entry per window')
        if not any(p.is_cuda for p in self.parameters()):
            raise _lib.CoskadError('the model must live on the CUDA device of traj (there is no CPU path)')
        ctx = self._context(traj)
        self._sync_encoder(ctx)
        rows = self._head()[2]
        Z = torch.empty((N, rows), device=traj.device, dtype=torch.float32) if want_latent else None
        score = cen = None
        if flavour != _lib.SCORE_NONE:
            cen = (self.c if center is None else center).to(device=traj.device, dtype=torch.float32).contiguous().view(-1)
            if score_out is not None:
                assert score_out.is_cuda and score_out.dtype == torch.float32 and score_out.is_contiguous() \
                    and score_out.numel() == N
                score = score_out
            else:
                score = torch.empty((N,), device=traj.device, dtype=torch.float32)
        rc = ctx.lib.coskad_encode_score_traj_fwd(ctx.h, int(flavour), traj.data_ptr(), traj.shape[0], win_row.data_ptr(),
                                                  _lib._ptr(trans), _lib._ptr(mats), n_mats, _lib._ptr(cen), N,
                                                  _lib._ptr(Z), _lib._ptr(score), _lib.stream_ptr(traj.device))
        ctx.check(rc, 'coskad_encode_score_traj_fwd')
        return Z, score


class STSAE(STSE):
    """models/sts/ae.py:168-264 -- encoder + rev_btlnk + STS-GCN decoder; forward returns (Z, X_hat)."""

    def build_model(self) -> None:
        super().build_model()
        self.rev_btlnk = nn.Linear(in_features=self.latent_dim,
                                   out_features=self.hidden_dimension * self.n_frames * self.n_joints)
        self._set_decoder_type()

    def _set_decoder_type(self) -> None:
        if self.encoder_type != 'sts_gcn':
            raise ValueError(f'No decoder available for encoder type {self.encoder_type}.')
        self.decoder = Decoder(self.input_dim, self.layer_channels, self.hidden_dimension, self.n_frames,
                               self.n_joints, self.dropout, self.bias)

    def _sync_decoder(self, ctx: _lib.Context) -> None:
        tensors = list(self.decoder.parameters()) + list(self.decoder.buffers()) + [self.rev_btlnk.weight, self.rev_btlnk.bias]
        key = self._version_key(tensors)
        if key == self._dec_key:
            return
        arr = self.decoder.layer_params_array()
        rc = ctx.lib.coskad_set_decoder(ctx.h, self.rev_btlnk.weight.data_ptr(), self.rev_btlnk.bias.data_ptr(),
                                        self.latent_dim, len(self.decoder.model), arr,
                                        _lib.stream_ptr(self.rev_btlnk.weight.device))
        ctx.check(rc, 'coskad_set_decoder')
        self._dec_key = key

    @torch.no_grad()
    def autoencode_score(self, X: torch.Tensor, center: Optional[torch.Tensor] = None, want_xhat: bool = True,
                         want_scores: bool = True):
        """Fused eval path of the auto-encoder: (Z, X_hat, rec_score, lat_score) from one kernel."""
        X = _check_x(X, self.input_dim, self.n_frames, self.n_joints)
        ctx = self._context(X)
        self._sync_encoder(ctx)
        self._sync_decoder(ctx)
        B = X.shape[0]
        Z = torch.empty((B, self.latent_dim), device=X.device, dtype=torch.float32)
        Xh = torch.empty_like(X) if want_xhat else None
        rec = torch.empty((B,), device=X.device, dtype=torch.float32) if want_scores else None
        lat = torch.empty((B,), device=X.device, dtype=torch.float32) if want_scores else None
        cen = (self.c if center is None else center).to(device=X.device, dtype=torch.float32).contiguous().view(-1)
        rc = ctx.lib.coskad_autoencode_score_fwd(ctx.h, X.data_ptr(), cen.data_ptr(), B, Z.data_ptr(), _lib._ptr(Xh),
                                                 _lib._ptr(rec), _lib._ptr(lat), _lib.stream_ptr(X.device))
        ctx.check(rc, 'coskad_autoencode_score_fwd')
        return Z, Xh, rec, lat

    def decode(self, Z: torch.Tensor, input_shape) -> torch.Tensor:
        from .train import decode_forward
        return decode_forward(self, Z)

    def forward(self, X: torch.Tensor):
        if self.training:
            from .train import stsae_train_forward
            return stsae_train_forward(self, X)
        Z, Xh, _, _ = self.autoencode_score(X, want_scores=False)
        return Z, Xh
