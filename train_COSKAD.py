#!/usr/bin/env python
"""train_COSKAD.py -- same CLI as the reference entry point (train_COSKAD.py:18-85):

    python train_COSKAD.py --config config/UBnormal/hyperbolic_encoder.yaml
    torchrun --nproc-per-node 8 train_COSKAD.py --config ...        # data parallel, one rank per B200

The LightningModule is picked from the flags use_decoder / use_vae / hyperbolic / static_center, the Lightning
Trainer + DDPStrategy is replaced by coskad_b200.trainer.Trainer (flat NCCL gradient all-reduce, all-reduced
center).  With ``dataset_choice: synthetic`` (or --synthetic) a synthetic dataset in the reference's batch-tuple
format is used; otherwise the reference's own loader (utils.dataset.get_dataset_and_loader) must be importable.
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from coskad_b200 import config as ccfg          # noqa: E402
from coskad_b200 import tasks                   # noqa: E402
from coskad_b200.trainer import Trainer         # noqa: E402


def get_loaders(args, ae_args):
    if getattr(args, 'dataset_choice', '') == 'synthetic':
        from coskad_b200.data import get_dataset_and_loader
    else:
        from utils.dataset import get_dataset_and_loader          # the reference's CPU data prep (out of scope here)
    ds, loader = get_dataset_and_loader(ae_args, split='train')
    val = None
    if getattr(args, 'validation', False):
        vds, val = get_dataset_and_loader(ae_args, split='validation' if args.dataset_choice != 'synthetic' else 'test',
                                          validation=True)
        if args.dataset_choice == 'synthetic':
            args.gt_table = (vds.clips, vds.gts)
    return loader, val


def main(argv=None):
    parser = argparse.ArgumentParser(description='Pose_AD_Experiment')
    parser.add_argument('-c', '--config', type=str, required=True)
    parser.add_argument('--synthetic', action='store_true', help='use the synthetic stand-in dataset')
    parser.add_argument('--epochs', type=int, default=None)
    parser.add_argument('--cuda_graph', action='store_true',
                        help='replay the training step as one CUDA graph (also: cuda_graph: true in the YAML)')
    cli = parser.parse_args(argv)
    args = ccfg.load_config(cli.config)
    if cli.synthetic:
        args.dataset_choice = 'synthetic'
    if cli.epochs is not None:
        args.ae_epochs = cli.epochs
    args, ae_args, _dcec, _res, _opt = ccfg.init_sub_args(args)
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        torch.distributed.init_process_group('nccl', device_id=torch.device('cuda', local))
    Litmodel = tasks.select_task(args)
    model = Litmodel(args)
    loader, val = get_loaders(args, ae_args)
    monitor = 'validation_auc' if getattr(args, 'validation', False) else 'loss'
    trainer = Trainer(max_epochs=args.ae_epochs, device=torch.device('cuda', local), ckpt_dir=args.ckpt_dir,
                      monitor=monitor, mode='max' if monitor == 'validation_auc' else 'min', save_top_k=2,
                      cuda_graph=cli.cuda_graph or bool(getattr(args, 'cuda_graph', False)))
    trainer.fit(model, loader, val)
    if world > 1:
        torch.distributed.destroy_process_group()
    return trainer


if __name__ == '__main__':
    main()
