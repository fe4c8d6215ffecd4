"""Flat YAML -> Namespace handling with the reference's prefix convention (utils/argparser.py:10-45,154-166):
keys prefixed ``dataset_``, ``ae_``, ``res_``, ``opt_`` are also exposed, prefix stripped, in sub-namespaces;
``debug: True`` forces 10 epochs; experiment directories are created at parse time."""
from __future__ import annotations

import argparse
import os
from typing import Tuple

import yaml

DEFAULTS = dict(debug=False, seed=999, device='cuda', accelerator='gpu', devices=[0], num_coords=2, h_dim=64, latent_dim=16,
                channels=[32, 16, 32], dropout=0.0, projector='linear', distance='euclidean', encoder_type='sts_gcn',
                use_decoder=False, use_vae=False, hyperbolic=False, static_center=False, distribution='ps',
                alpha=1e-6, lambda_=0.01, beta=0.0, gamma=0.0, phi=1.0, center_tolerance=1e-3, pad_size=-1, smoothing=50, validation=False,
                use_hr=False, ae_epochs=100, opt_lr=1e-4, dataset_choice='UBnormal', dataset_seg_len=12, dataset_batch_size=2048,
                dataset_num_transform=5, dataset_headless=False, dataset_kp18_format=False, dataset_double_item=False,
                exp_dir='./checkpoints', dir_name='coskad_b200_run', load_ckpt='', split='test', gt_path='', wandb=False)


def sub_namespace(args: argparse.Namespace, prefix: str) -> argparse.Namespace:
    """utils/argparser.py:154-166: keys starting with ``prefix`` with the prefix removed"""
    return argparse.Namespace(**{k[len(prefix):]: v for k, v in vars(args).items() if k.startswith(prefix)})


def init_sub_args(args: argparse.Namespace, make_dirs: bool = True) -> Tuple[argparse.Namespace, ...]:
    """(args, data_args(=ae_args upstream), dcec/res args, res_args, opt_args) like utils/argparser.py:10-45"""
    for k, v in DEFAULTS.items():
        if not hasattr(args, k):
            setattr(args, k, v)
    if args.debug:
        args.ae_epochs = 10                      # utils/argparser.py:11-12
    if make_dirs:
        args.ckpt_dir = os.path.join(args.exp_dir, args.dataset_choice, args.dir_name)
        os.makedirs(args.ckpt_dir, exist_ok=True)
    ae_args = sub_namespace(args, 'dataset_')
    for k in ('num_coords', 'seed', 'device', 'exp_dir', 'split'):
        setattr(ae_args, k, getattr(args, k))
    return args, ae_args, sub_namespace(args, 'ae_'), sub_namespace(args, 'res_'), sub_namespace(args, 'opt_')


def load_config(path: str) -> argparse.Namespace:
    with open(path) as f:
        return argparse.Namespace(**yaml.load(f, Loader=yaml.FullLoader))
