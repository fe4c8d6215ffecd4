"""Host logic of the minimal trainer (coskad_b200/trainer.py) that needs no GPU: lazy logging, and the optimizer /
scheduler plumbing the CUDA-graph replay of the training step relies on."""
import torch

from coskad_b200.trainer import LightningModule, _make_capturable, _same_layout, _flat_tensors


def test_log_is_lazy_and_reads_the_last_value():
    m = LightningModule()
    t = torch.tensor(1.5)
    m.log('loss', t)
    m.log('plain', 2)
    assert torch.is_tensor(m._log_raw['loss'])            # no conversion (= no device synchronisation) per step
    t.fill_(2.5)                                          # a replayed CUDA graph overwrites the logged scalar in place
    assert m._logged == {'loss': 2.5, 'plain': 2.0}


def test_capturable_lr_is_updated_in_place_by_both_schedulers():
    """hyperbolic_encoder.py:198-217: ReduceLROnPlateau(max, 0.2, min 1e-6) when validating, else CosineAnnealingLR; the
    captured optimizer step reads the learning rate from a device scalar, so the schedulers must write into that tensor"""
    p = torch.nn.Parameter(torch.ones(4))
    for kind in ('cosine', 'plateau'):
        opt = torch.optim.Adam([p], lr=1e-3, fused=True)
        if kind == 'cosine':
            sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=10, eta_min=1e-5)
        else:
            sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, mode='max', factor=0.2, patience=0, min_lr=1e-6)
        _make_capturable(opt, torch.device('cpu'))
        lr = opt.param_groups[0]['lr']
        assert torch.is_tensor(lr) and opt.param_groups[0]['capturable']
        for _ in range(3):
            p.grad = torch.ones(4)
            opt.step()
            sched.step(0.5) if kind == 'plateau' else sched.step()
        assert opt.param_groups[0]['lr'] is lr                      # same tensor, new value
        expected = 1e-5 + (1e-3 - 1e-5) * (1 + torch.cos(torch.tensor(torch.pi * 3 / 10))) / 2 if kind == 'cosine' else 1e-3 * 0.2 ** 2
        assert abs(float(lr) - float(expected)) < 1e-9, (kind, float(lr), float(expected))


def test_batch_layout_guard():
    a = [torch.zeros(8, 2, 12, 17), torch.zeros(8, dtype=torch.int64)]
    assert _flat_tensors(a) and _same_layout(a, [torch.empty_like(t) for t in a])
    assert not _same_layout([a[0][:5], a[1][:5]], a)                # ragged last batch -> eager step
    assert not _flat_tensors([a, a]) and not _flat_tensors([])     # nested batches (dataset_double_item) stay eager


def test_flat_adam_is_only_used_for_a_plain_cuda_adam_over_an_attached_bucket():
    """optim.FlatAdam.wrap declines (-> the caller keeps opt.step()) for CPU parameters, other optimizers, weight decay and
    unattached buckets: nothing of it runs without a GPU"""
    import torch
    from coskad_b200 import dist as cdist
    from coskad_b200.optim import FlatAdam
    ps = [torch.nn.Parameter(torch.randn(3, 4)), torch.nn.Parameter(torch.randn(5))]
    bucket = cdist.FlatGradBucket(ps)
    assert FlatAdam.wrap(torch.optim.Adam(ps, lr=1e-3), bucket) is None            # bucket not attached
    bucket.attach()
    assert bucket.flat.numel() % 4 == 0 and bucket.flat.numel() >= 17               # padded for the float4 kernel
    assert FlatAdam.wrap(torch.optim.Adam(ps, lr=1e-3), bucket) is None            # CPU parameters
    assert FlatAdam.wrap(torch.optim.SGD(ps, lr=1e-3), bucket) is None
    assert FlatAdam.wrap(None, bucket) is None


def test_regulariser_accumulates_into_marked_grad_views_like_autograd():
    """losses._L2Reg: with parameters marked by FlatGradBucket.attach() the regulariser's gradient goes into the bucket views with
    one multi-tensor add (and autograd gets None); unmarked parameters take the ordinary path.  Same numbers either way."""
    import torch
    from coskad_b200 import dist as cdist
    from coskad_b200.losses import calc_reg_loss

    def make():
        torch.manual_seed(3)
        m = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.BatchNorm1d(5), torch.nn.Linear(5, 2))
        return m

    a, b = make(), make()
    (calc_reg_loss(a) * 0.37).backward()
    bucket = cdist.FlatGradBucket(b.parameters()).attach()
    bucket.zero_()
    assert all(getattr(p, 'coskad_direct_grad', False) for p in b.parameters())
    (calc_reg_loss(b) * 0.37).backward()
    for (n, p), q in zip(a.named_parameters(), b.parameters()):
        if 'bias' in n:
            assert p.grad is None and float(q.grad.abs().sum()) == 0.0          # biases are not regularised (model_utils.py:92)
        else:
            assert torch.equal(p.grad, q.grad), n
        assert q.grad.data_ptr() >= bucket.flat.data_ptr()                      # still a view of the bucket
    bucket.detach()
    assert all(p.grad is None and not p.coskad_direct_grad for p in b.parameters())
