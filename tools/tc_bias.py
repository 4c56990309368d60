"""Is the tcgen05 engine's residual error a systematic gain (the tensor core's fp32 accumulator truncating toward zero) or noise?
For conv layers of growing K: least-squares gain g = <tc, simt> / <simt, simt> - 1 of the 3xTF32 result against the exact fp32 SIMT
engine, the normalised max error before and after removing that gain, and the number of accumulate steps per accumulator.
(GPU box; python tools/tc_bias.py)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # tools/ sits next to deft_b200/
sys.path.insert(0, ROOT)
import torch
from deft_b200 import ops, packing

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
print("kind        cin->cout k  nkb  steps/acc |   gain(tc3x)   nerr     nerr(gain removed) |  gain(all-positive data)")
for cin, cout, k, H, W in ((32, 64, 1, 64, 96), (64, 64, 3, 64, 96), (128, 128, 3, 48, 64), (256, 256, 3, 38, 68), (512, 256, 3, 19, 34), (512, 512, 3, 19, 34), (1280, 512, 1, 19, 34)):
    res = {}
    for name, pos in (("rand", False), ("pos", True)):
        w = torch.rand(cout, cin, k, k, generator=g) * 2 - 1
        x = torch.randn(1, cin, H, W, generator=g)
        if pos:
            w, x = w.abs(), x.abs()
        w = w * (3.0 / (cin * k * k)) ** 0.5
        wp, ldw = packing.pack_conv(w.to(dev))
        sc, sh = torch.ones(cout, device=dev), torch.zeros(cout, device=dev)
        xs = [ops.as_nhwc(x.to(dev))]
        outs = {}
        for label, flags in (("simt", 0), ("tc3x", 2)):
            outs[label] = ops.conv2d_bn_act(xs, wp, ldw, sc, sh, cout, k, k, 1, k // 2, False, flags=flags).nchw_view().double()
        torch.cuda.synchronize()
        a, b = outs["tc3x"], outs["simt"]
        gain = float((a * b).sum() / (b * b).sum()) - 1.0
        nerr = float((a - b).abs().max() / b.abs().max())
        nerr2 = float((a / (1.0 + gain) - b).abs().max() / b.abs().max())
        res[name] = (gain, nerr, nerr2)
    nkb = cin * k * k // 32
    steps = 12 * nkb // 2
    print(f"conv {cin:5d}->{cout:<4d} {k}  {nkb:4d}  {steps:6d}   | {res['rand'][0]:+.3e}  {res['rand'][1]:.2e}  {res['rand'][2]:.2e}          | {res['pos'][0]:+.3e} (nerr {res['pos'][1]:.2e} -> {res['pos'][2]:.2e})", flush=True)
