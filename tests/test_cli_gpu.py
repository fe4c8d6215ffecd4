"""GPU: the reference's command lines end to end on the synthetic dataset -- train_COSKAD.py writes checkpoints, eval_COSKAD.py
loads the best one and reports the same kind of AUC the reference prints (train_COSKAD.py:18-85, eval_COSKAD.py:49-253)."""
import glob
import os
import shutil

import pytest
import yaml

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('cfg', ['hyperbolic_encoder', 'hyperbolic_encoder_static', 'euclidean_encoder', 'euclidean_autoencoder',
                                 'spherical_vae'])
def test_train_then_eval_cli(tmp_path, cfg):
    import eval_COSKAD
    import train_COSKAD
    conf = yaml.safe_load(open(os.path.join(ROOT, 'config', 'synthetic', cfg + '.yaml')))
    conf['exp_dir'] = str(tmp_path)
    conf['ae_epochs'] = 2
    path = tmp_path / (cfg + '.yaml')
    path.write_text(yaml.safe_dump(conf))
    trainer = train_COSKAD.main(['-c', str(path)])
    ckpts = sorted(glob.glob(os.path.join(str(tmp_path), 'synthetic', conf['dir_name'], '*.ckpt')))
    assert ckpts, 'train_COSKAD.py wrote no checkpoint'
    conf['load_ckpt'] = os.path.basename(ckpts[-1])
    path.write_text(yaml.safe_dump(conf))
    auc = eval_COSKAD.main(['-c', str(path)])
    assert 0.0 <= auc <= 1.0
    shutil.rmtree(str(tmp_path), ignore_errors=True)
