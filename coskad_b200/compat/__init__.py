"""Drop-in shims for the reference's import paths.

``install()`` makes the reference's own task modules (``models/hyperbolic_encoder.py``, ``models/spherical_vae.py``,
``models/euclidean_autoencoder.py``, ``models/euclidean_encoder_staticCenter.py``, ``models/euclidean_encoder_dynamicCenter.py``)
importable and runnable on the CUDA hot path, unmodified.  It registers in ``sys.modules``:
  * ``models.stse.stse_hidden_hypersphere.STSE``, ``models.stsae.stsae_hidden_hypersphere.STSAE``,
    ``models.stsve.stsve_hidden_hypersphere.STSVE`` -- the three modules those files import but the reference does
    not ship (models/hyperbolic_encoder.py:16, models/euclidean_autoencoder.py:18, models/spherical_vae.py:16), with
    their OLD keyword names (c_in, h_dim, channels) mapped onto the in-tree names (input_dim, hidden_dimension,
    layer_channels) and the return orders those callers expect.  ``models`` itself is the reference's (namespace)
    package when the reference is on ``sys.path`` -- it is extended, never replaced -- and an empty package otherwise;
  * ``geoopt.manifolds.stereographic.math`` -> ``coskad_b200.gmath`` when geoopt is not installed;
  * ``pytorch_lightning`` (``LightningModule``, ``LightningDataModule``, ``Trainer``, ``callbacks.ModelCheckpoint``,
    ``loggers.WandbLogger``, ``strategies.DDPStrategy``) -> the minimal trainer when Lightning is not installed;
  * ``power_spherical.distributions`` -> the restated distributions of ``coskad_b200.spherical`` and a no-op
    ``matplotlib.pyplot`` (utils/eval_utils.py:5 imports it for one plotting helper) when those are not installed.
Nothing is overridden if the real package is importable.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types
from typing import Optional

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


def _module(name: str, package: bool = False, **attrs) -> types.ModuleType:
    mod = types.ModuleType(name)
    if package:
        mod.__path__ = []          # a package: `import name.sub` resolves through sys.modules
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    parent, _, leaf = name.rpartition('.')
    if parent and parent in sys.modules:
        setattr(sys.modules[parent], leaf, mod)
    return mod


def _have(name: str) -> bool:
    if name in sys.modules:
        return not getattr(sys.modules[name], '__coskad_shim__', False)
    try:
        return importlib.util.find_spec(name) is not None
    except (ImportError, ValueError):
        return False


def _shim(name: str, package: bool = False, **attrs) -> types.ModuleType:
    return _module(name, package, __coskad_shim__=True, **attrs)


def reference_root() -> Optional[str]:
    """the unmodified reference tree shipped beside the repo (baseline/_ref, git-ignored; oracle/install_ref.py)"""
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), 'baseline', '_ref')
    return root if os.path.isdir(os.path.join(root, 'models')) else None


def _models_package() -> types.ModuleType:
    """the reference's own ``models`` package when it is importable (extended in place), else an empty package"""
    if 'models' in sys.modules and hasattr(sys.modules['models'], '__path__'):
        return sys.modules['models']
    sys.modules.pop('models', None)            # a bare non-package module of that name cannot host sub-modules
    try:
        pk = importlib.import_module('models')
        if hasattr(pk, '__path__'):
            return pk
        sys.modules.pop('models', None)
    except ImportError:
        pass
    return _shim('models', package=True)


class _NullPlot(types.ModuleType):
    """matplotlib.pyplot stand-in: every attribute is a no-op callable (utils/eval_utils.py:221-228 only plots)"""

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return lambda *a, **k: None


class LightningDataModule:
    """pl.LightningDataModule as far as the reference's ``LitDataModule`` classes use it
    (models/hyperbolic_encoder.py:309-319): ``save_hyperparameters`` + ``hparams``"""

    def __init__(self) -> None:
        self.hparams = types.SimpleNamespace()

    def save_hyperparameters(self, *a, **k) -> None:
        import inspect
        frame = inspect.currentframe().f_back
        for name, val in frame.f_locals.items():
            if name not in ('self', '__class__'):
                setattr(self.hparams, name, val)


class ModelCheckpoint:
    """pytorch_lightning.callbacks.ModelCheckpoint(dirpath, save_top_k, monitor, mode) of train_COSKAD.py:70-73; the
    minimal Trainer reads these four fields"""

    def __init__(self, dirpath=None, save_top_k=1, monitor=None, mode='min', **_ignored) -> None:
        self.dirpath, self.save_top_k, self.monitor, self.mode = dirpath, save_top_k, monitor, mode


class DDPStrategy:
    """placeholder of train_COSKAD.py:78: data parallelism here is one process per GPU under torchrun"""

    def __init__(self, *a, **k) -> None:
        pass


class WandbLogger:
    def __init__(self, *a, **k) -> None:
        raise RuntimeError('wandb logging is out of scope of coskad_b200 (set use_wandb: False)')


def install(force: bool = False, reference: Optional[str] = None) -> None:
    """register the shims; ``reference`` (default: baseline/_ref when present) is appended to ``sys.path`` so that
    ``import models.hyperbolic_encoder`` finds the reference's own file"""
    ref = reference if reference is not None else reference_root()
    if ref and ref not in sys.path:
        sys.path.append(ref)
    if force or not _have('geoopt'):
        _shim('geoopt', package=True)
        _shim('geoopt.manifolds', package=True)
        _shim('geoopt.manifolds.stereographic', package=True)
        sys.modules['geoopt.manifolds.stereographic.math'] = _gmath
        sys.modules['geoopt.manifolds.stereographic'].math = _gmath
    if force or not _have('pytorch_lightning'):
        _shim('pytorch_lightning', package=True, LightningModule=_trainer.LightningModule,
              LightningDataModule=LightningDataModule, Trainer=_trainer.Trainer)
        _shim('pytorch_lightning.callbacks', ModelCheckpoint=ModelCheckpoint)
        _shim('pytorch_lightning.loggers', WandbLogger=WandbLogger)
        _shim('pytorch_lightning.strategies', DDPStrategy=DDPStrategy)
    if not _have('matplotlib'):
        _shim('matplotlib', package=True)
        plot = _NullPlot('matplotlib.pyplot')
        plot.__coskad_shim__ = True
        sys.modules['matplotlib.pyplot'] = plot
        sys.modules['matplotlib'].pyplot = plot
    from .. import spherical as _sph
    if not _have('power_spherical'):
        _shim('power_spherical', package=True, PowerSpherical=_sph.PowerSphericalQ,
              HypersphericalUniform=_sph.HypersphericalUniformP)
        _shim('power_spherical.distributions', PowerSpherical=_sph.PowerSphericalQ,
              HypersphericalUniform=_sph.HypersphericalUniformP)
    _models_package()
    for sub, cls_name, cls in (('stse', 'STSE', STSE), ('stsae', 'STSAE', STSAE), ('stsve', 'STSVE', _sph.STSVE)):
        _shim(f'models.{sub}', package=True)
        _shim(f'models.{sub}.{sub}_hidden_hypersphere', **{cls_name: cls})
