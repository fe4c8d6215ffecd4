"""Drop-in shims for the reference's import paths.

``install()`` registers, in ``sys.modules``:
  * ``models.stse.stse_hidden_hypersphere.STSE``, ``models.stsae.stsae_hidden_hypersphere.STSAE``,
    ``models.stsve.stsve_hidden_hypersphere.STSVE`` -- the three modules the reference's Lightning modules import
    but does not ship (models/hyperbolic_encoder.py:16, models/euclidean_autoencoder.py:18,
    models/spherical_vae.py:16), with their OLD keyword names (c_in, h_dim, channels) mapped onto the in-tree
    names (input_dim, hidden_dimension, layer_channels) and the return orders those callers expect;
  * ``geoopt.manifolds.stereographic.math`` -> ``coskad_b200.gmath`` when geoopt is not installed;
  * ``pytorch_lightning`` -> the minimal trainer when Lightning is not installed.
Nothing is overridden if the real package is importable.
"""
from __future__ import annotations

import importlib
import sys
import types

from .. import gmath as _gmath
from .. import sts as _sts
from .. import trainer as _trainer


class STSE(_sts.STSE):
    """old-kwarg constructor used at models/hyperbolic_encoder.py:61-63, euclidean_encoder_*.py:77-80"""

    def __init__(self, c_in, h_dim, latent_dim, n_frames, dropout, n_joints, channels, projector='linear',
                 encoder_type='sts_gcn', distance='euclidean', **kw):
        super().__init__(input_dim=c_in, layer_channels=list(channels), hidden_dimension=h_dim, latent_dim=latent_dim,
                         n_frames=n_frames, n_joints=n_joints, encoder_type=str(encoder_type).lower(), projector=projector,
                         distance=distance, dropout=dropout, **kw)


class STSAE(_sts.STSAE):
    """returns (X_hat, Z) like the module the reference imports (models/euclidean_autoencoder.py:111-115);
    the in-tree models/sts/ae.py:250 returns (Z, X_hat)."""

    def __init__(self, c_in, h_dim, latent_dim, n_frames, dropout, n_joints, channels, **kw):
        super().__init__(input_dim=c_in, layer_channels=list(channels), hidden_dimension=h_dim, latent_dim=latent_dim,
                         n_frames=n_frames, n_joints=n_joints, encoder_type='sts_gcn', projector='linear',
                         distance='euclidean', dropout=dropout)

    def forward(self, X):
        Z, Xh = super().forward(X)
        return Xh, Z


def _module(name: str, **attrs) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def _have(name: str) -> bool:
    try:
        importlib.import_module(name)
        return True
    except Exception:
        return False


def install(force: bool = False) -> None:
    pk = sys.modules.get('models') or _module('models')
    for sub, cls_name, cls in (('stse', 'STSE', STSE), ('stsae', 'STSAE', STSAE)):
        parent = _module(f'models.{sub}')
        leaf = _module(f'models.{sub}.{sub}_hidden_hypersphere', **{cls_name: cls})
        setattr(parent, f'{sub}_hidden_hypersphere', leaf)
        setattr(pk, sub, parent)
    try:
        from ..spherical import STSVE
        parent = _module('models.stsve')
        leaf = _module('models.stsve.stsve_hidden_hypersphere', STSVE=STSVE)
        parent.stsve_hidden_hypersphere = leaf
        pk.stsve = parent
    except ImportError:
        pass
    if force or not _have('geoopt'):
        geo = _module('geoopt')
        man = _module('geoopt.manifolds')
        ste = _module('geoopt.manifolds.stereographic')
        sys.modules['geoopt.manifolds.stereographic.math'] = _gmath
        geo.manifolds, man.stereographic, ste.math = man, ste, _gmath
    if force or not _have('pytorch_lightning'):
        _module('pytorch_lightning', LightningModule=_trainer.LightningModule, Trainer=_trainer.Trainer)
