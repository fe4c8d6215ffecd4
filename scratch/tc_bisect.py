import sys, os, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import torch
    from coskad_b200 import _lib
    stage = int(sys.argv[1])
    ctx = _lib.context(0)
    K = N = 16
    A = torch.randn(128, K, device='cuda'); Bh = torch.randn(N * K, device='cuda'); Bl = torch.zeros(N * K, device='cuda')
    out = torch.zeros(128, N, device='cuda')
    rc = ctx.lib.coskad_debug_tc_mix(ctx.h, A.data_ptr(), Bh.data_ptr(), Bl.data_ptr(), K, N, stage << 4, out.data_ptr(), 0)
    torch.cuda.synchronize()
    print('stage', stage, 'ok, out abs max', float(out.abs().max()))
else:
    for st in (1, 2, 3, 4):
        r = subprocess.run([sys.executable, __file__, str(st)], capture_output=True, text=True, timeout=120)
        print((r.stdout.strip().splitlines() or ['-'])[-1], '|', (r.stderr.strip().splitlines() or [''])[-1][:150])
