"""Training path: autograd Functions over the per-layer CUDA kernels (csrc/train_ops.cuh).

Reference semantics reproduced (autograd of the module graph):
    ST_GCNN_layer.forward          models/graph_layers/stsgcn.py:94-116   (train-mode BatchNorm2d, PReLU)
    Encoder / Decoder              models/common/components.py:94-105,168-179
    STSE.encode / STSAE.decode     models/sts/ae.py:76-105,210-230
BatchNorm uses per-GPU batch statistics (the reference has no SyncBN) and updates the running
statistics in place exactly like nn.BatchNorm2d (momentum 0.1, unbiased variance, num_batches_tracked).
The eval-mode variant of the same per-layer kernels (running statistics) serves the decode-only path.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib

P = 12 * 17
#: 1: the 1x1 convolutions (forward, backward data, weight gradient) run on the tensor cores (tcgen05, 3xTF32) with the
#: BatchNorm / PReLU backward fused into the backward kernels; 0: the FP32 CUDA-core kernels (A/B measurement)
TRAIN_IMPL = 1


def set_train_impl(impl: int) -> None:
    global TRAIN_IMPL
    if impl not in (0, 1):
        raise ValueError('train impl must be 0 (FP32 CUDA cores) or 1 (tcgen05)')
    TRAIN_IMPL = int(impl)


def _ctx(t: torch.Tensor) -> _lib.Context:
    c = _lib.context(t.device.index if t.device.index is not None else torch.cuda.current_device())
    if getattr(c, 'train_impl', 1) != TRAIN_IMPL:
        c.check(c.lib.coskad_set_train_impl(c.h, TRAIN_IMPL), 'coskad_set_train_impl')
        c.train_impl = TRAIN_IMPL
    return c


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


def _direct(p) -> Optional[torch.Tensor]:
    """the gradient destination the backward kernels may accumulate into themselves: ``p.grad`` when it is a view of an
    attached ``dist.FlatGradBucket`` (which zeroes it at the start of every step and marks the parameter).  Every
    second-stage reduction of the training kernels ADDS to its destination, so writing there is what autograd's
    AccumulateGrad would do -- minus one ``add_`` launch per parameter per step (~60 one-microsecond kernels).  The
    Function then returns ``None`` for that input.  Parameters that are not marked (plain ``loss.backward()`` without a
    bucket, ``torch.autograd.grad``) take the ordinary path: gradients are returned to autograd."""
    if p is None or not getattr(p, 'coskad_direct_grad', False):
        return None
    g = p.grad
    if g is None or g.dtype != torch.float32 or not g.is_cuda or not g.is_contiguous() or g.shape != p.shape:
        return None
    return g


class _LayerFn(torch.autograd.Function):
    """one ST_GCNN_layer, train (batch statistics) or eval (running statistics) BatchNorm"""

    @staticmethod
    def forward(ctx, X, A, T, W1, b1, g1, be1, W2, b2, g2, be2, slope, layer, training):
        X = _f32c(X)
        B, CI = X.shape[0], X.shape[1]
        CO = W1.shape[0]
        c = _ctx(X)
        st = _lib.stream_ptr(X.device)
        lib = c.lib
        G1 = torch.empty_like(X)
        G = torch.empty_like(X)
        c.check(lib.coskad_train_contract_fwd(c.h, X.data_ptr(), A.data_ptr(), T.data_ptr(), B * CI, G1.data_ptr(),
                                              G.data_ptr(), st), 'coskad_train_contract_fwd')
        y1 = torch.empty((B, CO, X.shape[2], X.shape[3]), device=X.device, dtype=torch.float32)
        y2 = torch.empty_like(y1)
        bn1, bn2 = layer.tcn[1], layer.residual[1]
        if training and getattr(c, 'train_impl', 1) == 1:
            # tensor-core path: convolutions + statistics, then ONE launch for the statistics' second stage + BatchNorm finalize
            mi = torch.empty(4 * CO, device=X.device, dtype=torch.float32)
            c.check(lib.coskad_train_mix_fwd_bn(c.h, G.data_ptr(), X.data_ptr(), W1.data_ptr(), _lib._ptr(b1), W2.data_ptr(),
                                                _lib._ptr(b2), B, CI, CO, y1.data_ptr(), y2.data_ptr(), float(bn1.eps),
                                                float(bn1.momentum), bn1.running_mean.data_ptr(), bn1.running_var.data_ptr(),
                                                bn2.running_mean.data_ptr(), bn2.running_var.data_ptr(), mi.data_ptr(),
                                                _nbt_ptr(bn1, X.device), _nbt_ptr(bn2, X.device), st), 'coskad_train_mix_fwd_bn')
        else:
            stats = torch.zeros(4 * CO, device=X.device, dtype=torch.float64)
            c.check(lib.coskad_train_mix_fwd(c.h, G.data_ptr(), X.data_ptr(), W1.data_ptr(), _lib._ptr(b1), W2.data_ptr(),
                                             _lib._ptr(b2), B, CI, CO, y1.data_ptr(), y2.data_ptr(), stats.data_ptr(), st),
                    'coskad_train_mix_fwd')
            if training:
                mi = torch.empty(4 * CO, device=X.device, dtype=torch.float32)
                c.check(lib.coskad_train_bn_finalize(c.h, stats.data_ptr(), B * P, CO, float(bn1.eps), float(bn1.momentum),
                                                     bn1.running_mean.data_ptr(), bn1.running_var.data_ptr(),
                                                     bn2.running_mean.data_ptr(), bn2.running_var.data_ptr(), mi.data_ptr(),
                                                     _nbt_ptr(bn1, X.device), _nbt_ptr(bn2, X.device), st),
                        'coskad_train_bn_finalize')
            else:   # eval: normalise with the running statistics (parameter prep, 4*CO numbers)
                mi = torch.cat([bn1.running_mean, torch.rsqrt(bn1.running_var + bn1.eps),
                                bn2.running_mean, torch.rsqrt(bn2.running_var + bn2.eps)]).to(torch.float32).contiguous()
        out = torch.empty_like(y1)
        c.check(lib.coskad_train_bn_prelu_fwd(c.h, y1.data_ptr(), y2.data_ptr(), mi.data_ptr(), g1.data_ptr(),
                                              be1.data_ptr(), g2.data_ptr(), be2.data_ptr(), slope.data_ptr(), B, CO,
                                              out.data_ptr(), st), 'coskad_train_bn_prelu_fwd')
        ctx.save_for_backward(X, G1, G, y1, y2, mi, A, T, W1, W2, g1, be1, g2, be2, slope)
        ctx.has_b = (b1 is not None, b2 is not None)
        ctx.params = (A, T, W1, b1, g1, be1, W2, b2, g2, be2, slope)      # the leaves themselves: backward looks at their .grad
        ctx.training = training
        return out

    @staticmethod
    def backward(ctx, dout):
        X, G1, G, y1, y2, mi, A, T, W1, W2, g1, be1, g2, be2, slope = ctx.saved_tensors
        if not ctx.training:
            raise RuntimeError('backward through an eval-mode ST_GCNN layer is not implemented (train() the model)')
        dout = _f32c(dout)
        B, CI, CO = X.shape[0], X.shape[1], W1.shape[0]
        c = _ctx(X)
        st = _lib.stream_ptr(X.device)
        lib = c.lib
        # Gradient destinations: the parameter's .grad view of the flat bucket when there is one (the kernels' second stages
        # add to their destination), else a slice of one zeroed buffer (one memset per layer) that is returned to autograd.
        pA, pT, pW1, pb1, pg1, pbe1, pW2, pb2, pg2, pbe2, pslope = ctx.params
        dests = [_direct(p) for p in (pW1, pW2, pb1, pb2, pA, pT, pg1, pbe1, pg2, pbe2, pslope)]
        # dA and dT take one second-stage launch when they are adjacent (as in the bucket): both direct or both temporary
        if dests[4] is None or dests[5] is None:
            dests[4] = dests[5] = None
        shapes = (W1.shape, W2.shape, (CO,), (CO,), A.shape, T.shape, (CO,), (CO,), (CO,), (CO,), (1,))
        need = [(i, int(torch.Size(sh).numel())) for i, sh in enumerate(shapes) if dests[i] is None]
        tc = getattr(c, 'train_impl', 1) == 1
        nred = (2 * (3 * CO + 1) + 3) // 4 * 4
        # tensor-core path: `red` is assigned by its kernel (coskad_train_bn_prelu_bwd_grads), so it needs no zero fill -- with every
        # gradient going straight into the bucket views the layer's backward then starts without a memset at all
        offs = [0 if tc else nred]
        for _, n in need:
            offs.append(offs[-1] + (n + 3) // 4 * 4)
        zbuf = torch.zeros(offs[-1], device=X.device, dtype=torch.float32) if offs[-1] else None
        red = (torch.empty(nred, device=X.device, dtype=torch.float32) if tc else zbuf[:nred]).view(torch.float64)[:3 * CO + 1]
        out = list(dests)
        for k, (i, n) in enumerate(need):
            out[i] = zbuf[offs[k]:offs[k] + n].view(shapes[i])
        dW1, dW2, db1, db2, dA, dT, dg1, dbe1, dg2, dbe2, dslope = out
        dy1 = torch.empty_like(y1)
        dy2 = torch.empty_like(y2)
        dG = torch.empty_like(X)
        dXres = torch.empty_like(X)
        if tc:      # reductions, then `red` + the BatchNorm / PReLU parameter gradients in one second-stage launch
            c.check(lib.coskad_train_bn_prelu_bwd_grads(c.h, dout.data_ptr(), y1.data_ptr(), y2.data_ptr(), mi.data_ptr(),
                                                        g1.data_ptr(), be1.data_ptr(), g2.data_ptr(), be2.data_ptr(),
                                                        slope.data_ptr(), B, CO, red.data_ptr(), dg1.data_ptr(), dbe1.data_ptr(),
                                                        dg2.data_ptr(), dbe2.data_ptr(), dslope.data_ptr(), st),
                    'coskad_train_bn_prelu_bwd_grads')
        else:
            c.check(lib.coskad_train_bn_prelu_bwd(c.h, dout.data_ptr(), y1.data_ptr(), y2.data_ptr(), mi.data_ptr(),
                                                  g1.data_ptr(), be1.data_ptr(), g2.data_ptr(), be2.data_ptr(), slope.data_ptr(),
                                                  B, CO, red.data_ptr(), dy1.data_ptr(), dy2.data_ptr(), st),
                    'coskad_train_bn_prelu_bwd')
            c.check(lib.coskad_train_bn_param_grads(c.h, red.data_ptr(), CO, dg1.data_ptr(), dbe1.data_ptr(), dg2.data_ptr(),
                                                    dbe2.data_ptr(), dslope.data_ptr(), st), 'coskad_train_bn_param_grads')
        if tc:      # BatchNorm / PReLU backward applied inside the tensor-core data-gradient kernel
            c.check(lib.coskad_train_mix_bwd_tc(c.h, dout.data_ptr(), y1.data_ptr(), y2.data_ptr(), mi.data_ptr(), g1.data_ptr(),
                                                be1.data_ptr(), g2.data_ptr(), be2.data_ptr(), slope.data_ptr(), red.data_ptr(),
                                                G.data_ptr(), X.data_ptr(), W1.data_ptr(), W2.data_ptr(), B, CI, CO,
                                                dy1.data_ptr(), dy2.data_ptr(), dG.data_ptr(), dXres.data_ptr(), dW1.data_ptr(),
                                                db1.data_ptr(), dW2.data_ptr(), db2.data_ptr(), st), 'coskad_train_mix_bwd_tc')
        else:
            c.check(lib.coskad_train_mix_bwd(c.h, dy1.data_ptr(), dy2.data_ptr(), G.data_ptr(), X.data_ptr(), W1.data_ptr(),
                                             W2.data_ptr(), B, CI, CO, dG.data_ptr(), dXres.data_ptr(), dW1.data_ptr(),
                                             db1.data_ptr(), dW2.data_ptr(), db2.data_ptr(), st), 'coskad_train_mix_bwd')
        dX = torch.empty_like(X)
        c.check(lib.coskad_train_contract_bwd(c.h, dG.data_ptr(), dXres.data_ptr(), X.data_ptr(), G1.data_ptr(),
                                              A.data_ptr(), T.data_ptr(), B * CI, dX.data_ptr(), dA.data_ptr(),
                                              dT.data_ptr(), st), 'coskad_train_contract_bwd')

        def ret(i, present=True):       # None: the kernels already accumulated into .grad (or the input does not exist)
            return out[i] if (present and dests[i] is None) else None
        return (dX, ret(4), ret(5), ret(0), ret(2, ctx.has_b[0]), ret(6), ret(7),
                ret(1), ret(3, ctx.has_b[1]), ret(8), ret(9), ret(10), None, None)


def _nbt_ptr(bn, device):
    """device pointer of ``num_batches_tracked`` (int64 scalar) or None: the finalize kernel increments it"""
    t = getattr(bn, 'num_batches_tracked', None)
    if t is None:
        return None
    if t.device != device or t.dtype != torch.int64:
        raise _lib.CoskadError(f'num_batches_tracked must be an int64 tensor on {device}, got {t.dtype} on {t.device}')
    return t.data_ptr()


def _check_params(*tensors) -> None:
    """the kernels read raw device pointers: a half / double / CPU / non-contiguous parameter would be read as garbage"""
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
            raise _lib.CoskadError('training kernels need contiguous float32 CUDA parameters, got '
                                   f'{tuple(t.shape)} {t.dtype} on {t.device} (contiguous={t.is_contiguous()})')


def layer_forward(layer, X: torch.Tensor, training: bool) -> torch.Tensor:
    if isinstance(layer.residual, torch.nn.Identity):
        raise NotImplementedError('identity residual (c_in == c_out) does not occur in any COSKAD config; not implemented '
                                  'in the training kernels')
    conv1, bn1, conv2, bn2 = layer.tcn[0], layer.tcn[1], layer.residual[0], layer.residual[1]
    _check_params(layer.gcn.A, layer.gcn.T, conv1.weight, conv1.bias, bn1.weight, bn1.bias, bn1.running_mean, bn1.running_var,
                  conv2.weight, conv2.bias, bn2.weight, bn2.bias, bn2.running_mean, bn2.running_var, layer.prelu.weight)
    return _LayerFn.apply(X, layer.gcn.A, layer.gcn.T, conv1.weight, conv1.bias, bn1.weight, bn1.bias,
                          conv2.weight, conv2.bias, bn2.weight, bn2.bias, layer.prelu.weight, layer, training)


class _LinearReduceFn(torch.autograd.Function):
    """out[B,D] = H[B,F] W^T + b  with W [D,F]  (btlnk / fc_mean / fc_var, models/sts/ae.py:155-157)"""

    @staticmethod
    def forward(ctx, H, W, bias):
        H = _f32c(H)
        B, F = H.shape
        D = W.shape[0]
        c = _ctx(H)
        out = torch.empty((B, D), device=H.device, dtype=torch.float32)
        c.check(c.lib.coskad_train_linear(c.h, 0, None, H.data_ptr(), W.data_ptr(), 0, _lib._ptr(bias), B, F, D,
                                          out.data_ptr(), _lib.stream_ptr(H.device)), 'coskad_train_linear(0)')
        ctx.save_for_backward(H, W)
        ctx.has_bias = bias is not None
        ctx.params = (W, bias)
        return out

    @staticmethod
    def backward(ctx, dz):
        H, W = ctx.saved_tensors
        dz = _f32c(dz)
        B, F = H.shape
        D = W.shape[0]
        c = _ctx(H)
        st = _lib.stream_ptr(H.device)
        dH = torch.empty_like(H)
        c.check(c.lib.coskad_train_linear(c.h, 1, dz.data_ptr(), None, W.data_ptr(), 0, None, B, F, D, dH.data_ptr(), st),
                'coskad_train_linear(1)')
        pW, pb = ctx.params
        gW, gb = _direct(pW), _direct(pb)
        dW = gW if gW is not None else torch.zeros_like(W)
        c.check(c.lib.coskad_train_linear(c.h, 2, dz.data_ptr(), H.data_ptr(), None, 0, None, B, F, D, dW.data_ptr(), st),
                'coskad_train_linear(2)')
        db = None
        if ctx.has_bias:
            db = gb if gb is not None else torch.zeros(D, device=H.device, dtype=torch.float32)
            c.check(c.lib.coskad_train_col_sum(c.h, dz.data_ptr(), B, D, db.data_ptr(), st), 'coskad_train_col_sum')
        return dH, (None if gW is not None else dW), (None if gb is not None else db)


class _LinearExpandFn(torch.autograd.Function):
    """out[B,F] = z[B,D] W^T + b  with W [F,D]  (rev_btlnk, models/sts/ae.py:206,222)"""

    @staticmethod
    def forward(ctx, z, W, bias):
        z = _f32c(z)
        B, D = z.shape
        F = W.shape[0]
        c = _ctx(z)
        out = torch.empty((B, F), device=z.device, dtype=torch.float32)
        c.check(c.lib.coskad_train_linear(c.h, 1, z.data_ptr(), None, W.data_ptr(), 1, _lib._ptr(bias), B, F, D,
                                          out.data_ptr(), _lib.stream_ptr(z.device)), 'coskad_train_linear(1)')
        ctx.save_for_backward(z, W)
        ctx.has_bias = bias is not None
        ctx.params = (W, bias)
        return out

    @staticmethod
    def backward(ctx, dH):
        z, W = ctx.saved_tensors
        dH = _f32c(dH)
        B, D = z.shape
        F = W.shape[0]
        c = _ctx(z)
        st = _lib.stream_ptr(z.device)
        dz = torch.empty_like(z)
        c.check(c.lib.coskad_train_linear(c.h, 0, None, dH.data_ptr(), W.data_ptr(), 1, None, B, F, D, dz.data_ptr(), st),
                'coskad_train_linear(0)')
        pW, pb = ctx.params
        gW, gb = _direct(pW), _direct(pb)
        dW = gW if gW is not None else torch.zeros_like(W)
        c.check(c.lib.coskad_train_linear(c.h, 2, z.data_ptr(), dH.data_ptr(), None, 1, None, B, F, D, dW.data_ptr(), st),
                'coskad_train_linear(2)')
        db = None
        if ctx.has_bias:
            db = gb if gb is not None else torch.zeros(F, device=z.device, dtype=torch.float32)
            c.check(c.lib.coskad_train_col_sum(c.h, dH.data_ptr(), B, F, db.data_ptr(), st), 'coskad_train_col_sum')
        return dz, (None if gW is not None else dW), (None if gb is not None else db)


def linear_reduce(H, W, bias):
    _check_params(W, bias)
    return _LinearReduceFn.apply(H, W, bias)


def linear_expand(z, W, bias):
    _check_params(W, bias)
    return _LinearExpandFn.apply(z, W, bias)


def stack_forward(stack, X: torch.Tensor, training: bool) -> torch.Tensor:
    for layer in stack.model:
        X = layer_forward(layer, X, training)
    return X


def encoder_features(model, X: torch.Tensor, training: bool = True) -> torch.Tensor:
    """flattened (c,t,v) encoder output [B, F] (models/sts/ae.py:88-100)"""
    H = stack_forward(model.encoder, X, training)
    return H.reshape(H.shape[0], -1)


def stse_train_forward(model, X: torch.Tensor) -> torch.Tensor:
    return linear_reduce(encoder_features(model, X, True), model.btlnk.weight, model.btlnk.bias)


def decode_forward(model, Z: torch.Tensor, training: Optional[bool] = None) -> torch.Tensor:
    """STSAE.decode (models/sts/ae.py:210-230): rev_btlnk, view [B,H,T,V], decoder stack"""
    training = model.training if training is None else training
    H = linear_expand(Z, model.rev_btlnk.weight, model.rev_btlnk.bias)
    H = H.view(Z.shape[0], model.hidden_dimension, model.n_frames, model.n_joints)
    return stack_forward(model.decoder, H, training)


def stsae_train_forward(model, X: torch.Tensor):
    Z = stse_train_forward(model, X)
    return Z, decode_forward(model, Z, True)
