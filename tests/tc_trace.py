"""Protocol timeline of the tcgen05 kernel (GPU box; perf bring-up): CTA 0 stamps clock64() at each protocol point."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from deft_b200 import ops, packing

dev = torch.device("cuda:0")
trace = torch.zeros(4 * 96 * 4, dtype=torch.int64, device=dev)
cin, cout, k, H, W = 64, 64, 3, 152, 272
w = torch.randn(cout, cin, k, k) * 0.05
wp, ldw = packing.pack_conv(w.to(dev))
sc, sh = torch.ones(cout, device=dev), torch.zeros(cout, device=dev)
x = ops.as_nhwc(torch.randn(1, cin, H, W).to(dev))
out = ops.NHWC.empty(1, H, W, cout, dev)
M = H * W
w_tc, bn = ops.tc_pack(wp, ldw, wp.shape[0], M, taps=k * k)
fn = lambda: ops.conv2d_bn_act([x], wp, ldw, sc, sh, cout, k, k, 1, k // 2, True, out=out, flags=2, w_tc=w_tc, tc_bn=bn)
for _ in range(3):
    fn()
torch.cuda.synchronize()
os.environ["DEFT_TC_TRACE"] = str(trace.data_ptr())
fn()
torch.cuda.synchronize()
t = trace.cpu().view(4, 96, 4)
t0 = int(t[t > 0].min())
names = ["MMA: loop_top  full_a_ok  full_b_ok  committed", "PROD: top  issued  empty_ok  arrived", "BCOPY: top  empty_ok", "EPI: top  acc_full_ok  done"]
for role in range(4):
    print(names[role])
    n = 44 if role < 3 else 24
    for i in range(n):
        row = t[role, i]
        if int(row.max()) == 0:
            continue
        print(f"  {i:3d} " + " ".join(f"{(int(v) - t0) if int(v) > 0 else -1:8d}" for v in row))
