"""training-step timing: hyperbolic dynamic-center step (fwd + bwd + Adam), batch 2048 per GPU (BASELINE configs[4])

    python tools/train_bench.py [B]                              one GPU, eager step + torch-profiler kernel table
    COSKAD_TB_GRAPH=1 python tools/train_bench.py                the step as one CUDA-graph replay (what Trainer(cuda_graph=True) does)
    torchrun --nproc-per-node N tools/train_bench.py             data parallel: flat NCCL gradient all-reduce per step, max over ranks
    COSKAD_TB_NOAR=1 / COSKAD_TB_FOREACH=1                       A/B switches: skip the all-reduce / for-each instead of fused Adam
    COSKAD_TRAIN_IMPL=0                                          A/B: the FP32 CUDA-core convolution kernels instead of tcgen05
    COSKAD_TB_NODIRECT=1                                         A/B: gradients handed to autograd instead of accumulated into the bucket views
    COSKAD_NO_FLAT_ADAM=1                                        A/B: torch's fused multi-tensor Adam instead of the flat Adam kernel
"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from coskad_b200 import synth, gmath, _lib, train as _train
_train.set_train_impl(int(os.environ.get('COSKAD_TRAIN_IMPL', '1')))      # 1: tcgen05 convolutions (default), 0: FP32 CUDA-core kernels
from coskad_b200.losses import calc_reg_loss

import torch.distributed as dist
from coskad_b200 import dist as cdist
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048      # windows per GPU per step (dataset_batch_size of the configs)
world, rank, local = int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('RANK', '0')), int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
    dist.init_process_group('nccl', device_id=dev)
m = synth.make_model('stse', 16, seed=0, device=dev).train()
cdist.broadcast_module_(m)
bucket = cdist.FlatGradBucket(m.parameters())
# COSKAD_TB_GRAPH=1: the whole step (fwd + bwd + Adam) as one CUDA graph replay; one process only (with the NCCL all-reduce
# inside the capture the 2-GPU replay hung on the test box)
GRAPH = bool(os.environ.get('COSKAD_TB_GRAPH')) and world == 1
opt = torch.optim.Adam(m.parameters(), lr=1e-4, capturable=GRAPH, fused=not os.environ.get('COSKAD_TB_FOREACH'))   # the tasks use fused=True
g = torch.Generator(device=dev).manual_seed(999 + rank)
x = torch.empty(B, 2, 12, 17, device=dev)
synth.synth_windows_(x, g)
c = torch.zeros(16, device=dev); c[0] = 0.1
acc = gmath.center_accumulator(16, dev)
DIRECT = not os.environ.get('COSKAD_TB_NODIRECT')         # A/B: COSKAD_TB_NODIRECT=1 = gradients through autograd's AccumulateGrad

flat_adam = None
if DIRECT and not os.environ.get('COSKAD_NO_FLAT_ADAM'):   # what trainer.TrainStep does: one flat Adam kernel instead of torch's multi-tensor step
    from coskad_b200.optim import FlatAdam
    bucket.attach()
    flat_adam = FlatAdam.wrap(opt, bucket)

def step():
    hidden = m(x)
    reg = calc_reg_loss(m)
    dist_c, hid = gmath.poincare_score(hidden, c, True)
    gmath.center_partial(hid, acc, _lib.SCORE_POINCARE)
    loss = dist_c.mean() + 1e-6 * reg
    if DIRECT: bucket.zero_()                             # attached bucket (trainer.TrainStep): kernels accumulate into the .grad views
    else: opt.zero_grad(set_to_none=True)
    loss.backward()
    if not os.environ.get('COSKAD_TB_NOAR'):
        bucket.allreduce_()                               # flat NCCL all-reduce of the gradients (no-op on one GPU)
    if flat_adam is not None: flat_adam.step()
    else: opt.step()
    return loss

if GRAPH:
    eager_step = step
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(3): eager_step()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        static_loss = eager_step()
    def step():
        graph.replay()
        return static_loss
for _ in range(5): step()
torch.cuda.synchronize()
if world > 1: dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
N = 20
t0 = time.perf_counter(); e0.record()
for _ in range(N): l = step()
e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
cdist.allreduce_center_acc(acc)                       # dynamic center: 18 doubles once per epoch
ms = e0.elapsed_time(e1) / N
if world > 1:
    t = torch.tensor([ms], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    if rank != 0:
        dist.destroy_process_group(); sys.exit(0)
    print(f'{world} GPUs x B={B}: {ms:.3f} ms/step (max over ranks; wall on rank 0 {(t1-t0)/N*1e3:.3f}) -> {world*B/ms*1e3:.0f} windows/s' + (' [no all-reduce]' if os.environ.get('COSKAD_TB_NOAR') else ''))
    dist.destroy_process_group(); sys.exit(0)
print(('[CUDA graph] ' if GRAPH else '') + f'B={B}: {ms:.3f} ms/step (device), {(t1-t0)/N*1e3:.3f} ms/step (wall) -> {B/ms*1e3:.0f} windows/s, loss {float(l):.4f}')
if GRAPH: sys.exit(0)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=45, max_name_column_width=60))
