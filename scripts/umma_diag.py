"""Diagnostic for the tcgen05 conv kernel on a real B200: prints error statistics for both
descriptor conventions (LBO/SBO as coded vs swapped). Not part of the product."""
import subprocess
import sys

CHILD = r'''
import sys, torch, torch.nn.functional as F
sys.path.insert(0, ".")
from deepbedmap_b200 import ops, _lib
swap = int(sys.argv[1])
_lib.call("dbm_debug_set", 1, swap)
torch.manual_seed(0)
for (n, cin, cs, h, w, cout) in [(1, 32, 4, 16, 16, 32), (1, 64, 8, 20, 23, 32), (2, 192, 24, 37, 41, 64)]:
    x = torch.randn(n, cs * 8, h, w).cuda().bfloat16().float()
    wt = (torch.randn(cout, cin, 3, 3) * 0.1).cuda().bfloat16().float()
    b = torch.randn(cout).cuda()
    s8 = ops.empty(n, cs, h, w, 8, dtype=torch.bfloat16)
    ops.nchw_to_slab8(x, s8)
    out32 = ops.empty(n, cout // 4, h, w, 4)
    ops.fill(out32, -777.0)
    ops.conv3x3_umma(s8, cin, ops.pack_conv3x3(wt, cout), b, cout, out_f32=out32)
    torch.cuda.synchronize()
    got = ops.slab4_to_nchw(out32, cout).double()
    ref = F.conv2d(x[:, :cin].double(), wt.double(), b.double(), padding=1)
    err = (got - ref).abs()
    print(f"swap={swap} case={(n, cin, cs, h, w, cout)} rel_l2={float((got-ref).norm()/ref.norm()):.3e} "
          f"max_abs={float(err.max()):.3e} frac_bad={float((err > 1e-3).double().mean()):.4f} "
          f"unwritten={int((got == -777.0).sum())}")
    if float(err.max()) > 1e-3:
        bad = (err > 1e-3).nonzero()[:8].tolist()
        print("   first bad (n,c,y,x):", bad)
        e0 = err[0]
        print("   bad per channel (first 8):", (e0 > 1e-3).flatten(1).sum(1)[:8].tolist())
        print("   bad per row y (first 20):", (e0 > 1e-3).sum((0, 2))[:20].tolist())
        print("   bad per col x (first 24):", (e0 > 1e-3).sum((0, 1))[:24].tolist())
'''

for swap in (0, 1):
    try:
        r = subprocess.run([sys.executable, "-c", CHILD, str(swap)], capture_output=True, text=True, timeout=240)
        print(r.stdout)
        if r.returncode:
            print(f"swap={swap}: exit {r.returncode}\n{r.stderr[-1500:]}")
    except subprocess.TimeoutExpired:
        print(f"swap={swap}: TIMEOUT")
