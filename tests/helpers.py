"""Shared helpers of the parity tests (oracle on CPU vs the CUDA path through the C ABI)."""
from __future__ import annotations

import numpy as np
import torch

from oracle import stsgcn as onet

CFG = dict(input_dim=2, layer_channels=[32, 16, 32], hidden_dimension=64, n_frames=12, n_joints=17)
CS = 205   # kCS of the fused kernel


def make_pair(kind: str = 'stse', latent_dim: int = 16, seed: int = 0):
    """(cuda module with the oracle's seeded state dict, that state dict on CPU)"""
    from coskad_b200 import sts
    sd = onet.init_state_dict(kind, latent_dim=latent_dim, seed=seed)
    cls = {'stse': sts.STSE, 'stsae': sts.STSAE}[kind]
    m = cls(latent_dim=latent_dim, encoder_type='sts_gcn', projector='linear', distance='euclidean', dropout=0.0, **CFG)
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval(), sd


def rel_err(a: torch.Tensor, b: torch.Tensor, atol: float = 0.0) -> float:
    a, b = a.double().cpu(), b.double().cpu()
    return float(((a - b).abs() / (b.abs() + atol + 1e-30)).max())


def stage_report(model, sd, x_cpu: torch.Tensor, with_decoder: bool = False) -> str:
    """Run the first tile of the fused kernel stage by stage and compare the CTA's shared-memory
    activations with the oracle's layer outputs; returns a text report (used in assert messages)."""
    from coskad_b200 import _lib
    lib = _lib.load()
    nw = lib.coskad_fused_tile_windows()
    nfl = lib.coskad_debug_fused_floats()
    x = x_cpu[:nw].contiguous()
    xd = x.cuda()
    model.encode_score(xd)           # makes sure the weights are packed
    ctx = model._ctx
    with torch.no_grad():
        _, acts = onet.layer_stack(x, sd, 'encoder', return_all=True)
        g1 = torch.einsum('nctv,vtq->ncqv', x, sd['encoder.model.0.gcn.T'])
        g = onet.graph_contract(x, sd['encoder.model.0.gcn.A'], sd['encoder.model.0.gcn.T'])
        z = onet.stse_forward(x, sd) if 'btlnk.weight' in sd else None
    R = nw * 32 * CS

    def rows(buf, off, C):   # [nw*C rows][205] -> [nw, C, 12, 17]
        return buf[off: off + nw * C * CS].view(nw * C, CS)[:, :204].reshape(nw, C, 12, 17)

    exp = {0: ('GB', 2 * R, 2, g1), 1: ('GB', 2 * R, 2, g), 2: ('R0', 0, 32, acts[0]), 5: ('R1', R, 16, acts[1]),
           8: ('R0', 0, 32, acts[2])}
    lines = []
    for stage in range(0, 13):
        out = torch.zeros(nfl, device='cuda')
        rc = lib.coskad_debug_fused_stage(ctx.h, int(with_decoder), xd.data_ptr(), x.shape[0], stage, out.data_ptr(),
                                          _lib.stream_ptr(xd.device))
        ctx.check(rc, 'coskad_debug_fused_stage')
        torch.cuda.synchronize()
        out = out.cpu()
        if stage in exp:
            name, off, C, ref = exp[stage]
            got = rows(out, off, C)
            lines.append(f'S{stage} {name}: max abs err {float((got - ref).abs().max()):.3e} (ref max {float(ref.abs().max()):.3e})')
        if stage == 12 and z is not None:
            got = out[2 * R + 4 * nw * 2 * CS // 2:]
            zf = out[-nw * 16:].view(nw, 16)[:, :z.shape[1]]
            lines.append(f'S12 zfin: max abs err {float((zf - z).abs().max()):.3e} (ref max {float(z.abs().max()):.3e})')
    return '\n'.join(lines)
