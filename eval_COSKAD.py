#!/usr/bin/env python
"""eval_COSKAD.py -- same CLI as the reference entry point (eval_COSKAD.py:47-253):

    python eval_COSKAD.py --config <exp_dir>/config.yaml

predict all windows (fused eval kernel), per-window score on the device, batched frame aggregation
(aggregate.score_and_aggregate) instead of the transformation x clip x person Python loops, then the same
score_process / roc_auc_score calls, per transformation and for the mean curve.
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from coskad_b200 import aggregate, config as ccfg, tasks      # noqa: E402
from coskad_b200.trainer import Trainer                       # noqa: E402


def evaluate(args, model, dataset_loader, ckpt_path=None, masks=None, device_tail=False):
    """returns (final AUC, per-transformation AUCs, curves).  device_tail=True (no pad_scores / HR masks): aggregation,
    smoothing and AUC all run on the GPU (aggregate.score_auc_device) and curves is None."""
    ds, loader = dataset_loader
    trainer = Trainer(device=torch.device('cuda', torch.cuda.current_device()), verbose=False)
    out = trainer.predict(model, dataloaders=loader, ckpt_path=ckpt_path, return_predictions=True)
    clips, gts = tasks.load_gt_table(args)
    nt = max(1, int(getattr(args, 'dataset_num_transform', 1)))
    pad = int(getattr(args, 'pad_size', -1))
    if args.use_decoder:
        o, hidden, gt_data, trans, meta, frames = tasks.light_processing_data(out)
        scores = model.window_scores(o, hidden, gt_data, loss_type='hyp')       # eval_COSKAD.py:66-73: rec_loss_weight = 0
    else:
        hidden, trans, meta, frames = tasks.light_processing_data(out)
        scores = model.window_scores(hidden, validation=False)
    if device_tail and pad == -1 and masks is None:
        auc, per_t = aggregate.score_auc_device(scores, trans, meta, frames, clips, nt, gts)
        return auc, per_t, None
    curves = aggregate.score_and_aggregate(scores, trans, meta, frames, clips, nt, pad_size=pad, gts=gts, masks=masks)
    auc, per_t = tasks.auc_from_curves(curves, clips, gts, masks)
    return auc, per_t, curves


def main(argv=None):
    parser = argparse.ArgumentParser(description='Pose_AD_Experiment')
    parser.add_argument('-c', '--config', type=str, required=True)
    parser.add_argument('--synthetic', action='store_true')
    cli = parser.parse_args(argv)
    args = ccfg.load_config(cli.config)
    if cli.synthetic:
        args.dataset_choice = 'synthetic'
    args, ae_args, _dcec, _res, _opt = ccfg.init_sub_args(args)
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    if getattr(args, 'dataset_choice', '') == 'synthetic':
        from coskad_b200.data import get_dataset_and_loader
    else:
        from utils.dataset import get_dataset_and_loader
    ds, loader = get_dataset_and_loader(ae_args, split=args.split)
    if args.dataset_choice == 'synthetic':
        args.gt_table = (ds.clips, ds.gts)
    model = tasks.select_task(args)(args)
    path = os.path.join(args.exp_dir, args.dataset_choice, args.dir_name, args.load_ckpt) if args.load_ckpt else None
    print('Loading model from {}'.format(path))
    # eval_COSKAD.py:92-101: HR subset of UBnormal = boolean frame masks per clip (the reference hard-codes their directory;
    # here it is the config key hr_masks_path, a glob of {scene}_{clip}.npy files)
    masks = tasks.hr_ubnormal(args.hr_masks_path) if getattr(args, 'use_hr', False) and getattr(args, 'hr_masks_path', '') else None
    auc, per_t, _ = evaluate(args, model, (ds, loader), ckpt_path=path, masks=masks)
    for t, a in per_t.items():
        print('auc = {} (transformation {})'.format(a, t + 1))
    print('final AUC score: {}'.format(auc))
    return auc


if __name__ == '__main__':
    main()
