#!/usr/bin/env python
"""bench.py -- headline benchmark of the COSKAD anomaly-scoring hot path on B200.

Metric (BASELINE.json): pose windows/sec scored.  Workload at N=1 (configs[1]): hyperbolic STS-GCN
encoder forward + Poincare distance scoring over 16 Mi synthetic UBnormal-shape windows resident in
HBM (27.4 GB), processed as `--steps` chunks of `--windows-per-step` windows (one step = one pass of
the fused kernel over one chunk; consecutive steps walk distinct chunks, each 1.7 GB >> L2).
Multi-GPU (torchrun, one rank per GPU): windows are sharded, every rank scores its own chunk per
step (weak scaling) and the scores are all-gathered over NCCL inside the timed region.

    python bench.py [--gpus N --steps K --warmup W] [--impl reference]

Prints ONE JSON line (rank 0).  Besides the headline it carries, measured in the same run under the same
clock sampler: `e2e` (pinned host windows -> scores on the host), `e2e_traj` (host trajectories -> scores),
`e2e_agg` (host trajectories + metadata -> scores -> frame aggregation -> smoothing -> AUC on the device),
`secondary` (BASELINE configs[2..4]: spherical VAE scoring, Euclidean auto-encoder scoring, the
data-parallel hyperbolic training step), `parity` (this run's GPU results against the CPU reference on
configs[0]), `parity_multi` (N > 1: a rank recomputes a neighbour's shard bit for bit), the rooflines
and the CPU legs.  `--impl reference` times the reference's own CPU path -- the unmodified
models/sts/ae.py STSE from baseline/_ref when that copy exists, else the oracle port -- on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_WINDOW = 3.947e6      # 1 973 496 MAC, reference operation order, D=16
FLOP_PER_WINDOW_VAE = 3.764e6  # encoder + fc_mean + fc_var, D=8
FLOP_PER_WINDOW_AE = 8.210e6   # encoder + rev_btlnk + decoder + reconstruction score, D=8
MIX_MAC_PER_WINDOW = 1279488   # channel mixing (tcn + residual) MACs per window: the tcgen05 share
BYTES_PER_WINDOW = 1636.0      # compulsory HBM bytes: x 1632 + score 4
TOTAL_WINDOWS = 16 * 1024 * 1024
TRAIN_BATCH = 2048             # dataset_batch_size of every reference config


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=16)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--windows-per-step', type=int, default=1024 * 1024)
    ap.add_argument('--resident-windows', type=int, default=TOTAL_WINDOWS,
                    help='windows kept resident in HBM per GPU (steps cycle through them)')
    ap.add_argument('--cpu-sample', type=int, default=4096, help='windows per CPU-baseline pass')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-secondary', action='store_true')
    ap.add_argument('--e2e-chunk', type=int, default=16384, help='windows per H2D chunk / kernel launch of the e2e leg')
    ap.add_argument('--agg-windows', type=int, default=0,
                    help='windows per GPU of the e2e_agg leg (0: 16 Mi on one GPU, 4 Mi per GPU otherwise)')
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '100', '-i', str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def window(self, t0: float, t1: float):
        sm, mx, reasons = [], [], set()
        for ts, line in list(self.rows):
            if ts < t0 - 0.05 or ts > t1 + 0.05:
                continue
            f = [v.strip() for v in line.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}

    def stop(self):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()


# ---- the CPU arm ----------------------------------------------------------------------------------
def reference_module():
    """the reference's OWN network, unmodified: models/sts/ae.py STSE imported from baseline/_ref (oracle/install_ref.py
    copies the tree there; it is git-ignored and travels to the GPU box).  None when the copy is absent."""
    ref = os.path.join(ROOT, 'baseline', '_ref')
    if not os.path.isdir(os.path.join(ref, 'models', 'sts')):
        return None
    if ref not in sys.path:
        sys.path.append(ref)
    try:
        import importlib
        return importlib.import_module('models.sts.ae').STSE
    except Exception:
        return None


def build_reference_net(RefSTSE, sd):
    """the reference's STSE at the config/UBnormal/hyperbolic_encoder.yaml shapes with the oracle's seeded state dict; its
    constructor prints ("Encoder type: ...", models/sts/ae.py): keep stdout for the one JSON line"""
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):
        net = RefSTSE(input_dim=2, layer_channels=[32, 16, 32], hidden_dimension=64, latent_dim=16, n_frames=12,
                      n_joints=17, encoder_type='sts_gcn', projector='linear', distance='euclidean', dropout=0.0)
    net.load_state_dict(sd, strict=True)
    return net.eval()


def cpu_path(sample: int, passes: int, warm: int = 1, min_seconds: float = 0.0, max_passes: int = 64):
    """The reference's CPU implementation of the path, all host threads: STSE forward (models/sts/ae.py:108-121) +
    project(expmap0) + dist (eval_COSKAD.py:194-196).  The network is the reference's own module when baseline/_ref is
    there (kind 'reference'), else the oracle port (kind 'port'); geoopt is not installable offline, so the 300-FLOP
    geometry tail is the restated geoopt math in both cases.  Returns (times, ncpu, nthreads, kind, scores)."""
    import torch
    from oracle import geoopt_math as ogm
    from oracle import stsgcn as onet
    ncpu = os.cpu_count() or 1
    torch.set_num_threads(ncpu)
    sd = onet.init_state_dict('stse', latent_dim=16, seed=0)
    x = onet.synth_windows(sample, seed=999)
    k = torch.tensor(-1.)
    c = torch.full((16,), 0.01)
    RefSTSE = reference_module()
    if RefSTSE is not None:
        net = build_reference_net(RefSTSE, sd)
        fwd, kind = (lambda xb: net(xb)), 'reference'
    else:
        fwd, kind = (lambda xb: onet.stse_forward(xb, sd)), 'port'
    times, scores = [], None
    with torch.no_grad():
        i = -1
        while True:
            i += 1
            if i >= warm + passes and (sum(times) >= min_seconds or len(times) >= max_passes):
                break
            t0 = time.perf_counter()
            out = []
            for lo in range(0, sample, 2048):                       # dataset_batch_size 2048
                z = fwd(x[lo:lo + 2048])
                out.append(ogm.dist(ogm.project(ogm.expmap0(z, k=k), k=k), c, k=k))
            dt = time.perf_counter() - t0
            if i >= warm:
                times.append(dt)
            scores = torch.cat(out)
    return times, ncpu, torch.get_num_threads(), kind, (x, sd, c, scores)


def cpu_aggregate_path(min_seconds: float = 2.0):
    """the reference's frame aggregation (utils/eval_utils.py:57-74 + eval_COSKAD.py:140-220, transcribed in
    oracle/aggregate.py) timed on the host: windows aggregated per second, scores given"""
    import numpy as np
    from oracle import aggregate as oagg
    trans, meta, frames, clips, gts = oagg.synth_dataset(n_clips=24, seed=3, num_transform=5, max_persons=5,
                                                         frame_range=(300, 600))
    scores = np.random.default_rng(0).random(len(trans)).astype(np.float32) + 0.1
    n, t = 0, 0.0
    while t < min_seconds and n < 8:
        t0 = time.perf_counter()
        oagg.aggregate_dataset(scores, trans, meta, frames, clips, 5)
        t += time.perf_counter() - t0
        n += 1
    return {'value': len(trans) * n / t, 'unit': 'windows/s', 'cores': 1, 'kind': 'port',
            'sample': f'{len(trans)} windows, {len(clips)} clips x 5 transformations, {n} passes (python loops are single-threaded)'}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    times, ncpu, nthr, kind, _ = cpu_path(args.cpu_sample, args.steps, max(args.warmup, 1))
    total = sum(times)
    v = args.cpu_sample * len(times) / total
    line = {
        'impl': 'reference', 'metric': 'pose windows/sec scored', 'value': v, 'unit': 'windows/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / len(times), 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'hyperbolic STS-GCN encoder fwd + Poincare distance scoring, UBnormal-shape windows '
                               '(BASELINE configs[1]); CPU reference arm: each step = a bounded sample of '
                               f'{args.cpu_sample} windows in batches of 2048',
                   'network': 'models/sts/ae.py STSE, unmodified, from baseline/_ref' if kind == 'reference'
                              else 'oracle port (baseline/_ref absent)'},
        'cpu_baseline': {'value': v, 'unit': 'windows/s', 'cores': nthr, 'kind': kind,
                         'sample': f'{args.cpu_sample} windows x {len(times)} passes, host cpu_count {ncpu}'},
        'e2e': {'value': v, 'unit': 'windows/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))
    return 0


# ---- our arm --------------------------------------------------------------------------------------
class Env:
    """rank / device / collectives / timing helpers shared by the legs"""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.rank = int(os.environ.get('RANK', '0'))
        self.local = int(os.environ.get('LOCAL_RANK', '0'))
        if not torch.cuda.is_available():
            raise SystemExit('bench.py needs a B200: no CUDA device visible (there is no CPU fallback for the product path)')
        torch.cuda.set_device(self.local)
        self.dev = torch.device('cuda', self.local)
        if self.world > 1:
            # NCCL prints its banner ("NCCL version ...") on fd 1 when the communicator is created: point fd 1 at stderr
            # while the process group and its communicator come up, so that stdout carries nothing but the JSON line
            sys.stdout.flush()
            saved_fd = os.dup(1)
            os.dup2(2, 1)
            try:
                dist.init_process_group('nccl', device_id=self.dev)
                dist.barrier()
                torch.cuda.synchronize()
            finally:
                sys.stdout.flush()
                os.dup2(saved_fd, 1)
                os.close(saved_fd)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_ms(self, ms: float) -> float:
        if self.world == 1:
            return ms
        t = self.torch.tensor([ms], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, steps: int, warm: int):
        """W untimed + K timed calls of fn(i), bracketed by barrier + synchronize, CUDA events, max over ranks.
        Returns (total_ms, per-step ms on this rank, (wall t0, wall t1))"""
        torch = self.torch
        for i in range(warm):
            fn(i)
        self.barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        t0 = time.time()
        ev[0].record()
        for i in range(steps):
            fn(warm + i)
            ev[i + 1].record()
        self.barrier()
        t1 = time.time()
        total = self.max_ms(ev[0].elapsed_time(ev[-1]))
        return total, [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)], (t0, t1)


def leg_e2e_agg(env: Env, model, center, n_windows: int, n_steps: int):
    """host trajectories + per-window metadata -> scores -> frame aggregation -> smoothing -> AUC, everything after the
    H2D copies on the device: eval_COSKAD.py:107-253 for a synthetic dataset (num_transform 5, 300-frame tracks, 6
    persons per clip).  The CSR grouping is built on the device (aggregate.DeviceGroupIndex)."""
    import numpy as np
    torch = env.torch
    from coskad_b200 import _lib, aggregate
    from coskad_b200.pipeline import TrajectoryScorer
    n_tr, plen, ppc = 5, 300, 6
    per_person = plen - 12 + 1
    persons = max(ppc, (n_windows // (n_tr * per_person)) // ppc * ppc)
    n_clips = persons // ppc
    pid = np.arange(persons, dtype=np.int64)
    start = np.arange(per_person, dtype=np.int64)
    rows1 = (pid[:, None] * plen + start[None, :]).reshape(-1)                 # one transformation
    clip_of_p = pid // ppc
    meta1 = np.stack([np.repeat(1 + clip_of_p // 100, per_person), np.repeat(1 + clip_of_p % 100, per_person),
                      np.repeat(1 + pid % ppc, per_person), np.tile(start + 1, persons)], axis=1)
    frames1 = (np.tile(start + 1, persons)[:, None] + np.arange(12, dtype=np.int64)[None, :])
    n1 = rows1.shape[0]
    N = n1 * n_tr
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    rows_h = pin(np.tile(rows1, n_tr))
    trans_h = pin(np.repeat(np.arange(n_tr, dtype=np.int32), n1))
    meta_h = pin(np.tile(meta1, (n_tr, 1)))
    frames_h = pin(np.tile(frames1, (n_tr, 1)))
    g = torch.Generator().manual_seed(1234 + env.rank)
    traj_h = (torch.randn(persons * plen, 34, generator=g) * 0.4).clamp_(-3, 3).pin_memory()
    clips = [(1 + c // 100, 1 + c % 100, plen) for c in range(n_clips)]
    rng = np.random.default_rng(7)
    gts = {(s, c): (rng.random(plen) < 0.1).astype(np.int64) for s, c, _ in clips}
    import math
    c45 = math.cos(math.radians(45.0))
    mats = torch.tensor([[[1, 0, 0], [0, 1, 0]], [[-1, 0, 0], [0, 1, 0]], [[0, -1, 0], [1, 0, 0]],
                         [[0, 1, 0], [1, 0, 0]], [[c45, -c45, 0], [c45, c45, 0]]], dtype=torch.float32)
    ts = TrajectoryScorer(model, _lib.SCORE_POINCARE, device=env.local)
    side = torch.cuda.Stream(env.dev)
    result = {}

    def step(_i):
        comp = torch.cuda.current_stream(env.dev)
        with torch.cuda.stream(side):                                           # metadata rides beside the scoring
            meta_d = meta_h.to(env.dev, non_blocking=True)
            frames_d = frames_h.to(env.dev, non_blocking=True)
            trans_d = trans_h.to(env.dev, non_blocking=True)
        dscore = ts.score(traj_h, rows_h, trans_h, mats, center=center, keep_on_device=True)
        comp.wait_stream(side)
        auc, per_t = aggregate.score_auc_device(dscore, trans_d, meta_d, frames_d, clips, n_tr, gts)   # D2H: n_tr + 1 doubles
        for t_ in (meta_d, frames_d, trans_d):
            t_.record_stream(comp)
        result['auc'] = auc

    ts.h2d_bytes = 0
    total, _, _ = env.timed(step, n_steps, 1)
    h2d = ts.h2d_bytes // (n_steps + 1) + meta_h.numel() * 8 + frames_h.numel() * 8 + trans_h.numel() * 4
    return {'value': N * env.world * n_steps / (total * 1e-3), 'unit': 'windows/s', 'windows_per_step_per_gpu': N,
            'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': 8 * (n_tr + 1), 'steps': n_steps,
            'clips': n_clips * n_tr, 'persons': persons * n_tr, 'auc': result.get('auc'),
            'note': 'TrajectoryScorer (host trajectories + window rows + transform ids) -> device scores; meta [N,4] + frames '
                    '[N,12] int64 H2D on a side stream; DeviceGroupIndex (sort of packed ids on the device) -> '
                    'coskad_frame_aggregate (float64, bit-exact order) -> coskad_score_process (shift 11 + Gaussian sigma 30) '
                    '-> per-transformation + mean-curve AUC on the device; random labels, so the AUC is ~0.5 by construction'}


def leg_scoring(env: Env, kind: str, steps: int, warm: int, fp32_peak: float):
    """BASELINE configs[2] (spherical VAE: encoder + fc_mean/fc_var + cosine score, UBnormal shape) and configs[3]
    (Euclidean auto-encoder: encoder + decoder + reconstruction and latent scores, STC shape), window-sharded"""
    torch = env.torch
    from coskad_b200 import _lib
    from coskad_b200.synth import make_model, synth_windows_
    W, nch = 256 * 1024, 4                                   # 4 x 428 MB resident chunks >> L2
    g = torch.Generator(device=env.dev).manual_seed(4321 + env.rank)
    x = torch.empty((W * nch, 2, 12, 17), device=env.dev, dtype=torch.float32)
    gathered = torch.empty(W * env.world, device=env.dev) if env.world > 1 else None
    if kind == 'vae':
        model = make_model('stsvae', 8, seed=0, device=env.dev)
        synth_windows_(x, g, 'ubnormal')
        mv = torch.nn.functional.normalize(torch.randn(8, generator=torch.Generator().manual_seed(0)), dim=0).to(env.dev)
        score = torch.empty(W, device=env.dev)

        def step(i):
            xi = x[(i % nch) * W:((i % nch) + 1) * W]
            model.encode_score(xi, _lib.SCORE_COSINE, center=mv, want_latent=False, score_out=score)
            if env.world > 1:
                env.dist.all_gather_into_tensor(gathered, score)
        flop, what = FLOP_PER_WINDOW_VAE, ('spherical VAE (use_vae) encoder fwd + fc_mean/fc_var head + hypersphere (cosine) score '
                                           'of Z_mean, latent 8, UBnormal-shape windows (BASELINE configs[2])')
    else:
        model = make_model('stsae', 8, seed=0, device=env.dev)
        synth_windows_(x, g, 'stc')
        cen = torch.full((8,), 0.05, device=env.dev)

        def step(i):
            xi = x[(i % nch) * W:((i % nch) + 1) * W]
            _, _, rec, lat = model.autoencode_score(xi, center=cen, want_xhat=False)
            if env.world > 1:
                env.dist.all_gather_into_tensor(gathered, rec)
        flop, what = FLOP_PER_WINDOW_AE, ('Euclidean auto-encoder (use_decoder) fwd: encoder + decoder + reconstruction and '
                                          'latent-distance scores, latent 8, STC-shape windows (BASELINE configs[3])')
    ctx = model._ctx if model._ctx is not None else None
    total, per, _ = env.timed(step, steps, warm)
    ach = flop * W * env.world * steps / (total * 1e-3) / 1e12 / env.world
    del x
    torch.cuda.empty_cache()
    return {'value': W * env.world * steps / (total * 1e-3), 'unit': 'windows/s', 'ms_per_step': total / steps, 'steps': steps,
            'windows_per_step_per_gpu': W, 'workload': what,
            'roofline': {'bound': 'fp32_fma', 'achieved': ach, 'peak': fp32_peak, 'unit': 'TFLOP/s per GPU',
                         'frac': ach / fp32_peak if fp32_peak else None, 'flop_per_window': flop}}


def leg_train(env: Env, steps: int, warm: int, fp32_peak: float):
    """BASELINE configs[4]: hyperbolic dynamic-center training step, 2 048 windows per GPU: forward + Poincare loss +
    regulariser + backward + flat NCCL gradient all-reduce + Adam + center partial sums, replayed as CUDA graphs around
    the collective (trainer.TrainStep); the center all-reduce + finalisation of the epoch end is inside the timed region."""
    torch = env.torch
    from coskad_b200 import config as ccfg, dist as cdist, tasks
    from coskad_b200.synth import randomize_bn_, synth_windows_
    from coskad_b200.trainer import TrainStep, _make_capturable
    ns = argparse.Namespace(hyperbolic=True, static_center=False, use_decoder=False, use_vae=False, latent_dim=16,
                            dataset_batch_size=TRAIN_BATCH, projector='linear', ae_epochs=100, opt_lr=1e-4, validation=False)
    args, *_ = ccfg.init_sub_args(ns, make_dirs=False)
    torch.manual_seed(0)
    lit = tasks.LitEncoder(args)
    randomize_bn_(lit.model, 0)
    lit.to(env.dev)
    cdist.broadcast_module_(lit)
    nb = 8
    g = torch.Generator(device=env.dev).manual_seed(999 + env.rank)
    pool = torch.empty((nb * TRAIN_BATCH, 2, 12, 17), device=env.dev)
    synth_windows_(pool, g, 'ubnormal')
    aux = [torch.zeros(TRAIN_BATCH, dtype=torch.int64, device=env.dev), torch.zeros((TRAIN_BATCH, 4), dtype=torch.int64, device=env.dev),
           torch.ones((TRAIN_BATCH, 12), dtype=torch.int64, device=env.dev)]
    batches = [[pool[i * TRAIN_BATCH:(i + 1) * TRAIN_BATCH]] + aux for i in range(nb)]
    c0 = torch.zeros(16, device=env.dev)
    c0[0] = 0.05
    lit.model.c = c0
    lit.temp = c0.clone()
    lit.train()
    opt = lit.configure_optimizers()['optimizer']
    _make_capturable(opt, env.dev)
    bucket = cdist.FlatGradBucket(lit.parameters()).attach()
    ts = TrainStep(lit, opt, bucket, env.dev)
    lit.on_train_epoch_start()
    for i in range(3):
        ts.eager(batches[i], i)
    ts.capture(batches[3], 3)
    losses = []

    def step(i):
        losses.append(ts.replay(batches[i % nb]))

    for i in range(warm):
        step(i)
    # one untimed epoch end as well: its first call pays one-off costs (lazy CUDA module loads of the norm / logging ops, the NCCL
    # communicator's first small all-reduce) that made the timed region vary by up to 27 ms from run to run
    lit.training_epoch_end([])
    env.barrier()
    lit.on_train_epoch_start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    trace = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)] if os.environ.get('COSKAD_BENCH_TRAIN_TRACE') else None
    e0.record()
    for i in range(steps):
        if trace:
            trace[i].record()
        step(warm + i)
    if trace:
        trace[steps].record()
    lit.training_epoch_end([])                               # center: all-reduce of D+2 doubles + finalisation
    e1.record()
    env.barrier()
    total = env.max_ms(e0.elapsed_time(e1))
    if trace:            # debugging aid: per-step device times of the timed region
        print('train_step trace (ms):', ' '.join(f'{trace[i].elapsed_time(trace[i + 1]):.3f}' for i in range(steps)),
              '| epoch end', f'{trace[steps].elapsed_time(e1):.3f}', file=sys.stderr)
    loss = float(losses[-1])
    ach = 3 * FLOP_PER_WINDOW * TRAIN_BATCH * steps / (total * 1e-3) / 1e12
    del pool
    torch.cuda.empty_cache()
    return {'value': TRAIN_BATCH * env.world * steps / (total * 1e-3), 'unit': 'windows/s', 'ms_per_step': total / steps,
            'steps': steps, 'windows_per_step_per_gpu': TRAIN_BATCH, 'loss': loss, 'finite': bool(loss == loss and abs(loss) < 1e30),
            'graphs': 'one CUDA graph per step' if env.world == 1 else 'graph A (fwd+bwd into the flat bucket) -> in-place '
                      'ncclAllReduce -> graph B (Adam), one stream, no host sync',
            'allreduce_bytes_per_step': int(bucket.numel * 4) if env.world > 1 else 0,
            'workload': 'hyperbolic dynamic-center training step (fwd + Poincare loss + reg + bwd + flat gradient all-reduce + '
                        'Adam + center partial sums), 2 048 windows per GPU, data-parallel (BASELINE configs[4]); the working '
                        'set of a step (activations ~1.5 GB) exceeds L2',
            'roofline': {'bound': 'fp32_fma', 'achieved': ach, 'peak': fp32_peak, 'unit': 'TFLOP/s per GPU',
                         'frac': ach / fp32_peak if fp32_peak else None,
                         'flop_per_step': 3 * FLOP_PER_WINDOW * TRAIN_BATCH, 'note': 'algorithmic: 3 x forward FLOPs'}}


def leg_parity_multi(env: Env, model, center, x, W: int):
    """N > 1: every rank regenerates its right neighbour's first chunk (same seeded generator), scores it locally and
    compares the result BIT FOR BIT with the neighbour's segment of the all-gathered scores"""
    torch = env.torch
    from coskad_b200 import _lib
    from coskad_b200.synth import synth_windows_
    n = min(int(x.shape[0]), 1 << 20)          # the first chunk synth_windows_ generated: same call sizes -> same values
    s_loc = torch.empty(n, device=env.dev)
    model.encode_score(x[:n], _lib.SCORE_POINCARE, center=center, want_latent=False, score_out=s_loc)
    gathered = torch.empty(n * env.world, device=env.dev)
    env.dist.all_gather_into_tensor(gathered, s_loc)
    nb = (env.rank + 1) % env.world
    xn = torch.empty((n, 2, 12, 17), device=env.dev)
    synth_windows_(xn, torch.Generator(device=env.dev).manual_seed(999 + nb), 'ubnormal')
    s_nb = torch.empty(n, device=env.dev)
    model.encode_score(xn, _lib.SCORE_POINCARE, center=center, want_latent=False, score_out=s_nb)
    ok = torch.tensor([1.0 if torch.equal(s_nb, gathered[nb * n:(nb + 1) * n]) else 0.0], device=env.dev)
    env.dist.all_reduce(ok, op=env.dist.ReduceOp.MIN)
    return bool(ok.item() == 1.0)


def leg_cpu_and_parity(env: Env, args):
    """rank 0, N = 1: the CPU baseline (configs[0]) AND, on the very same windows and weights, the parity gates of
    BASELINE.md section 5 evaluated in this run: per-window score rtol (pure relative), frame aggregation bit-exact, AUC"""
    import numpy as np
    torch = env.torch
    from coskad_b200 import _lib, aggregate, sts
    from coskad_b200.data import SyntheticPoseDataset
    from oracle import aggregate as oagg
    from oracle import geoopt_math as ogm
    from oracle import stsgcn as onet
    times, ncpu, nthr, kind, (x, sd, c, s_cpu) = cpu_path(args.cpu_sample, 3, 1, min_seconds=10.0)      # ~10 s of CPU work
    cpu_baseline = {'value': args.cpu_sample * len(times) / sum(times), 'unit': 'windows/s', 'cores': nthr, 'kind': kind,
                    'sample': f'{args.cpu_sample} windows x {len(times)} passes (batches of 2048), host cpu_count {ncpu}; network = '
                              + ('the reference\'s own models/sts/ae.py STSE (baseline/_ref)' if kind == 'reference' else 'oracle port')}
    m = sts.STSE(input_dim=2, layer_channels=[32, 16, 32], hidden_dimension=64, latent_dim=16, n_frames=12, n_joints=17,
                 encoder_type='sts_gcn', projector='linear', distance='euclidean', dropout=0.0)
    m.load_state_dict(sd, strict=True)
    m = m.to(env.dev).eval()
    parity = {'tolerance': 'per-window score |gpu - cpu| <= 1e-4 |cpu| (pure relative, fp32)', 'windows': int(args.cpu_sample)}
    k = torch.tensor(-1.)
    for shape in ('ubnormal', 'stc'):
        xs = x if shape == 'ubnormal' else onet.synth_windows(args.cpu_sample, seed=999, shape='stc')
        if shape == 'ubnormal':
            ref = s_cpu
        else:
            with torch.no_grad():
                ref = ogm.dist(ogm.project(ogm.expmap0(onet.stse_forward(xs, sd), k=k), k=k), c, k=k)
        _, s_gpu = m.encode_score(xs.to(env.dev), _lib.SCORE_POINCARE, center=c.to(env.dev))
        rel = ((s_gpu.cpu().double() - ref.double()).abs() / ref.double().abs()).max()
        parity[f'score_max_rel_err_{shape}'] = float(rel)
    # aggregation + AUC on a synthetic dataset (6 clips, 2 transformations)
    ds = SyntheticPoseDataset(n_clips=6, seed=5, num_transform=2)
    with torch.no_grad():
        s_ref = ogm.dist(ogm.project(ogm.expmap0(onet.stse_forward(ds.x, sd), k=k), k=k), c, k=k)
    _, s_dev = m.encode_score(ds.x.to(env.dev), _lib.SCORE_POINCARE, center=c.to(env.dev))
    ref_curves = oagg.aggregate_dataset(s_ref.numpy(), ds.trans.numpy(), ds.meta.numpy(), ds.frames.numpy(), ds.clips, 2)
    our_curves = aggregate.score_and_aggregate(s_ref.to(env.dev), ds.trans, ds.meta, ds.frames, ds.clips, 2)
    parity['aggregation_bit_exact'] = all(np.array_equal(a, b) for t in range(2) for a, b in zip(ref_curves[t], our_curves[t]))
    from sklearn.metrics import roc_auc_score
    gt = np.concatenate([ds.gts[(s_, c_)] for s_, c_, _ in ds.clips])
    auc_ref = float(roc_auc_score(gt, np.mean(np.stack([np.concatenate(ref_curves[t]) for t in range(2)], 0), 0)))
    auc_gpu, _ = aggregate.score_auc_device(s_dev, ds.trans, ds.meta, ds.frames, ds.clips, 2, ds.gts)
    parity.update(auc_cpu_reference=auc_ref, auc_gpu=auc_gpu, auc_equal_4_decimals=round(auc_ref, 4) == round(auc_gpu, 4))
    parity['ok'] = bool(max(parity['score_max_rel_err_ubnormal'], parity['score_max_rel_err_stc']) <= 1e-4
                        and parity['aggregation_bit_exact'] and parity['auc_equal_4_decimals'])
    # the reference's network under PyTorch eager on this GPU (context only: the reference ships no kernel of its own)
    ref_gpu = None
    RefSTSE = reference_module()
    if RefSTSE is not None:
        net = build_reference_net(RefSTSE, sd).to(env.dev)
        xb = x[:2048].to(env.dev)
        cd = c.to(env.dev)
        with torch.no_grad():
            f = lambda: ogm.dist(ogm.project(ogm.expmap0(net(xb), k=k), k=k), cd, k=k)
            for _ in range(3):
                f()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                f()
            e1.record()
            torch.cuda.synchronize()
        ref_gpu = {'value': 2048 * 20 / (e0.elapsed_time(e1) * 1e-3), 'unit': 'windows/s', 'batch': 2048,
                   'note': 'the reference\'s models/sts/ae.py STSE + restated geoopt scoring under PyTorch eager (cuDNN / cuBLAS '
                           'library kernels) on this B200, batch 2048 = dataset_batch_size; context only'}
    return cpu_baseline, parity, ref_gpu


def run_ours(args):
    env = Env()
    torch, dist = env.torch, env.dist
    world, rank, local, dev = env.world, env.rank, env.local, env.dev
    import ctypes
    from coskad_b200 import _lib, dist as cdist, gmath
    from coskad_b200.pipeline import HostScorer, TrajectoryScorer
    from coskad_b200.synth import make_model, synth_windows_

    model = make_model('stse', 16, seed=0, device=dev)     # random init, randomised BN statistics
    W = args.windows_per_step
    resident = max(W, min(args.resident_windows, TOTAL_WINDOWS))
    nchunks = resident // W
    # synthetic UBnormal-shape windows generated on the device, seed 999 + rank
    g = torch.Generator(device=dev).manual_seed(999 + rank)
    x = torch.empty((nchunks * W, 2, 12, 17), device=dev, dtype=torch.float32)
    synth_windows_(x, g, 'ubnormal')
    # center: gyro-midpoint of the first 65 536 projected embeddings of EVERY rank (float64 partial sums, all-reduced)
    z0, _ = model.encode_score(x[:65536])
    acc = gmath.center_accumulator(16, dev)
    gmath.center_partial(gmath.expmap0_project(z0), acc, _lib.SCORE_POINCARE)
    cdist.allreduce_center_acc(acc)
    center = gmath.center_finalize(acc, 16, _lib.SCORE_POINCARE)
    scores = torch.empty(W, device=dev, dtype=torch.float32)
    gathered = torch.empty(W * world, device=dev, dtype=torch.float32) if world > 1 else None
    ctx = model._ctx

    def step(i):
        xi = x[(i % nchunks) * W:((i % nchunks) + 1) * W]
        model.encode_score(xi, _lib.SCORE_POINCARE, center=center, want_latent=False, score_out=scores)
        if world > 1:
            dist.all_gather_into_tensor(gathered, scores)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = ctx.launch_count()
    for i in range(args.warmup):
        step(i)
    env.barrier()
    launches0 = ctx.launch_count()
    total_ms, kern_ms, (t_wall0, t_wall1) = env.timed(step, args.steps, 0)
    launches = ctx.launch_count() - launches0
    clocks = sampler.window(t_wall0, t_wall1) if rank == 0 else None
    value = W * world * args.steps / (total_ms * 1e-3)

    # ---- measured peaks of this GPU (FP32 FFMA loop, tcgen05 kind::tf32 loop) ----------------------------------------------
    tf = ctypes.c_double(0.0)
    ctx.check(ctx.lib.coskad_measure_fp32_peak(ctx.h, ctypes.byref(tf), _lib.stream_ptr(dev)), 'coskad_measure_fp32_peak')
    fp32_peak = float(tf.value)
    ctx.check(ctx.lib.coskad_measure_tf32_peak(ctx.h, ctypes.byref(tf), _lib.stream_ptr(dev)), 'coskad_measure_tf32_peak')
    tf32_peak = float(tf.value)

    # ---- e2e through the public host API: pinned host windows -> scores on the host ----------------
    e2e = e2e_traj = e2e_agg = None
    n_e2e = max(3, min(args.steps, 8))
    if not args.no_e2e:
        hs = HostScorer(model, _lib.SCORE_POINCARE, chunk=args.e2e_chunk, device=local)
        xh = torch.empty((W, 2, 12, 17), dtype=torch.float32).pin_memory()
        xh.copy_(x[:W])
        oh = torch.empty(W * world, dtype=torch.float32).pin_memory() if world > 1 else torch.empty(W, dtype=torch.float32).pin_memory()
        g_e2e = torch.empty(W * world, device=dev, dtype=torch.float32) if world > 1 else None

        def e2e_step(_i):
            if world > 1:      # the score gather of the sharded job is part of the end-to-end step
                hs.score(xh, oh[:W], center=center, on_device_scores=lambda d: dist.all_gather_into_tensor(g_e2e, d))
                oh.copy_(g_e2e, non_blocking=True)
                torch.cuda.current_stream(dev).synchronize()
            else:
                hs.score(xh, oh, center=center)
        e2e_step(0)
        hs.h2d_bytes = hs.d2h_bytes = 0
        t, _, _ = env.timed(e2e_step, n_e2e, 1)
        per = n_e2e + 1
        e2e = {'value': W * world * n_e2e / (t * 1e-3), 'unit': 'windows/s',
               'h2d_bytes_per_step': hs.h2d_bytes // per, 'd2h_bytes_per_step': hs.d2h_bytes // per + (W * world * 4 if world > 1 else 0),
               'steps': n_e2e, 'note': 'HostScorer: pinned host chunk -> H2D (copy stream) -> fused kernel -> '
                                       + ('NCCL all-gather of the scores -> ' if world > 1 else '') + 'D2H scores'}
        if rank == 0 and not bool(torch.isfinite(oh).all()):
            raise SystemExit('non-finite scores in the e2e pass')
        # ---- the same W windows scored from host TRAJECTORIES (window construction + 5 test-time transforms in-kernel)
        n_tr, plen = 5, 300                                    # num_transform of the UBnormal configs; frames per person
        per_person = plen - 12 + 1
        persons = max(1, W // (n_tr * per_person))
        base = torch.arange(per_person, dtype=torch.int64).repeat(persons) + \
            torch.arange(persons, dtype=torch.int64).repeat_interleave(per_person) * plen
        rows_h = base.repeat(n_tr).pin_memory()
        trans_h = torch.arange(n_tr, dtype=torch.int32).repeat_interleave(base.numel()).pin_memory()
        traj_h = torch.empty((persons * plen, 34), dtype=torch.float32).pin_memory()
        traj_h.copy_(x[: (persons * plen * 34 + 407) // 408].reshape(-1)[: persons * plen * 34].view(-1, 34))
        import math as _m
        c45 = _m.cos(_m.radians(45.0))
        mats = torch.tensor([[[1, 0, 0], [0, 1, 0]], [[-1, 0, 0], [0, 1, 0]], [[0, -1, 0], [1, 0, 0]],
                             [[0, 1, 0], [1, 0, 0]], [[c45, -c45, 0], [c45, c45, 0]]], dtype=torch.float32)
        ts = TrajectoryScorer(model, _lib.SCORE_POINCARE, device=local)
        oh2 = torch.empty(rows_h.numel(), dtype=torch.float32).pin_memory()
        ts.score(traj_h, rows_h, trans_h, mats, oh2, center=center)
        ts.h2d_bytes = ts.d2h_bytes = 0
        t, _, _ = env.timed(lambda _i: ts.score(traj_h, rows_h, trans_h, mats, oh2, center=center), n_e2e, 1)
        e2e_traj = {'value': rows_h.numel() * world * n_e2e / (t * 1e-3), 'unit': 'windows/s',
                    'h2d_bytes_per_step': ts.h2d_bytes // per, 'd2h_bytes_per_step': ts.d2h_bytes // per,
                    'windows_per_step_per_gpu': rows_h.numel(), 'steps': n_e2e,
                    'note': 'TrajectoryScorer: pinned host trajectory rows [frames, 34] + window start rows + transform ids -> '
                            'H2D -> fused kernel builds the stride-1 windows and applies the 5 test-time affine transforms '
                            'in its input stage -> D2H scores'}
        if rank == 0 and not bool(torch.isfinite(oh2).all()):
            raise SystemExit('non-finite scores in the trajectory e2e pass')
        del xh, oh, rows_h, trans_h, traj_h, oh2
        n_agg = args.agg_windows or (TOTAL_WINDOWS if world == 1 else 4 * 1024 * 1024)
        e2e_agg = leg_e2e_agg(env, model, center, n_agg, 2)

    # ---- BASELINE configs[2..4] in the same run -----------------------------------------------------
    secondary = None
    if not args.no_secondary:
        ns = max(3, min(args.steps, 8))
        t_s0 = time.time()
        secondary = {'vae': leg_scoring(env, 'vae', ns, 3, fp32_peak), 'ae': leg_scoring(env, 'ae', ns, 3, fp32_peak),
                     'train_step': leg_train(env, max(8, min(2 * args.steps, 32)), 5, fp32_peak)}
        if rank == 0:
            secondary['clocks'] = sampler.window(t_s0, time.time())
    parity_multi = leg_parity_multi(env, model, center, x, W) if world > 1 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    sampler.stop()

    # ---- roofline of the dominant kernel (the fused kernel is the only kernel of a step) ------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    fp32_peak_derived = 148 * 128 * 2 * float(peaks.get('sm_max_mhz', 1965.0)) * 1e6 / 1e12
    k_ms = statistics.mean(kern_ms) if world == 1 else total_ms / args.steps
    ach_tf = FLOP_PER_WINDOW * W / (k_ms * 1e-3) / 1e12
    ach_gb = BYTES_PER_WINDOW * W / (k_ms * 1e-3) / 1e9
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    # CONSTANT, not measured in this run: ncu dram__bytes_read+write of one profiled launch of this kernel
    # (profiles/r02_fused_eval_tc.md: 1726.40 + 10.27 MB for 1 048 576 windows; r01_v9: 1726.73 + 10.98 MB), scaled to W
    traffic_per_window = 1736.67e6 / 1048576
    l2_to_sm_per_window = 380.164e9 / 1048576      # l1tex__m_xbar2l1tex_read_bytes of the same capture
    roofline = {'bound': 'fp32_fma', 'achieved': ach_tf, 'peak': fp32_peak, 'unit': 'TFLOP/s',
                'frac': ach_tf / fp32_peak if fp32_peak else None,
                'traffic': traffic_per_window * W,
                'traffic_unit': 'bytes per launch; a CONSTANT from one ncu --set full capture (dram read+write of the profiled '
                                'launch, profiles/r02_fused_eval_tc.md), scaled to this launch -- not re-measured in this run',
                'l2_to_sm_bytes_per_window': l2_to_sm_per_window,
                'kernel': 'fused_eval_tc_kernel', 'launch_ms': k_ms,
                'peak_source': 'measured on this GPU by coskad_measure_fp32_peak (register-resident FFMA loop)',
                'peak_derived': fp32_peak_derived, 'frac_of_derived': ach_tf / fp32_peak_derived,
                'flop_per_window': FLOP_PER_WINDOW,
                'note': 'algorithmic FLOPs (3.947 MFLOP/window, SURVEY.md 8-d) over the launch time, against the FP32-FMA '
                        'peak: the fused path is compute bound (2 400 FLOP/B). 65 % of the MACs (channel mixing) execute on '
                        'tcgen05 kind::tf32 with 3xTF32 split operands, the graph contractions and the linear head on the '
                        'FP32 pipe (packed FFMA2), so this fraction mixes two pipes: roofline_fp32_pipe and roofline_tensor '
                        'split it per pipe; roofline_hbm gives the HBM view of the same launch'}
    fp32_flop = FLOP_PER_WINDOW - 2.0 * MIX_MAC_PER_WINDOW
    roofline_fp32_pipe = {'bound': 'fp32_fma', 'achieved': fp32_flop * W / (k_ms * 1e-3) / 1e12, 'peak': fp32_peak, 'unit': 'TFLOP/s',
                          'frac': fp32_flop * W / (k_ms * 1e-3) / 1e12 / fp32_peak if fp32_peak else None,
                          'flop_per_window': fp32_flop,
                          'note': 'only the work that runs on the FP32 pipe (graph contractions + head, 1.39 MFLOP/window) '
                                  'against the measured FFMA peak: the utilisation view the headline fraction hides'}
    roofline_hbm = {'bound': 'hbm', 'achieved': ach_gb, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': ach_gb / hbm_peak,
                    'traffic': traffic_per_window * W, 'bytes_per_window': BYTES_PER_WINDOW,
                    'peak_source': 'MEASURED_PEAKS.json hbm_gbs' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s'}
    mix_tf = 2.0 * MIX_MAC_PER_WINDOW * W / (k_ms * 1e-3) / 1e12
    roofline_tensor = {'bound': 'tensor', 'achieved': 3.0 * mix_tf, 'peak': tf32_peak, 'unit': 'TFLOP/s',
                       'frac': 3.0 * mix_tf / tf32_peak if tf32_peak else None, 'algorithmic_tflops': mix_tf,
                       'peak_source': 'measured on this GPU by coskad_measure_tf32_peak (back-to-back tcgen05 kind::tf32 '
                                      'M128 N256 K8 MMAs, A in TMEM, B in shared memory, all SMs)',
                       'bf16_cublas_peak_for_context': float(peaks.get('bf16_tflops_sustained', peaks.get('bf16_tflops', 1360.6))),
                       'note': 'the channel-mixing MACs (1.28 M per window) that run on tcgen05, counted three times (3xTF32 split: '
                               'hi*hi + lo*hi + hi*lo keeps the 1e-4 score tolerance), over the launch time, against the measured '
                               'TF32 tensor peak. The mixing GEMMs are K <= 64, N <= 64: the tensor pipe is a helper here, not the bound'}
    cpu_baseline = parity = ref_gpu = cpu_agg = None
    if not args.no_cpu_baseline and world == 1:
        cpu_baseline, parity, ref_gpu = leg_cpu_and_parity(env, args)
        cpu_agg = cpu_aggregate_path()

    line = {
        'metric': 'pose windows/sec scored', 'value': value, 'unit': 'windows/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': total_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'hyperbolic STS-GCN encoder fwd + Poincare distance scoring, UBnormal-shape windows '
                               '[2,12,17], channels 2-32-16-32-64, latent 16 (BASELINE configs[1])',
                   'windows_per_step_per_gpu': W, 'resident_windows_per_gpu': nchunks * W,
                   'l2_policy': 'inputs larger than L2: each step reads a distinct 1.7 GB chunk',
                   'parallelism': f'window-sharded x{world}' + (', NCCL all-gather of scores per step; center = all-reduced '
                                                                  'float64 partial sums' if world > 1 else '')},
        'e2e': e2e, 'e2e_traj': e2e_traj, 'e2e_agg': e2e_agg, 'gpu_launches': int(launches), 'clocks': clocks,
        'roofline': roofline, 'roofline_fp32_pipe': roofline_fp32_pipe, 'roofline_hbm': roofline_hbm,
        'roofline_tensor': roofline_tensor, 'secondary': secondary, 'parity': parity, 'parity_multi': parity_multi,
        'cpu_baseline': cpu_baseline, 'cpu_baseline_aggregate': cpu_agg, 'ref_gpu_eager': ref_gpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    a = parse()
    sys.exit(run_reference(a) if a.impl == 'reference' else run_ours(a))
