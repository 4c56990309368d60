"""One layer, a few launches (GPU box; for ncu --launch-skip/--launch-count captures): python tools/one_layer.py <case> [n]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # tools/ sits next to deft_b200/
sys.path.insert(0, ROOT)
import torch
from deft_b200 import ops, packing
from oracle import spec

CASES = {"conv64": ("conv", 64, 64, 3, 152, 272), "conv128": ("conv", 128, 128, 3, 76, 136), "conv32": ("conv", 64, 32, 3, 152, 272),
         "conv1x1": ("conv", 128, 64, 1, 152, 272), "dcn64": ("dcn", 64, 64, 3, 152, 272), "dcn128": ("dcn", 128, 128, 3, 76, 136),
         "dcn256": ("dcn", 256, 256, 3, 38, 68)}
case = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
kind, cin, cout, k, H, W = CASES[case]
dev = torch.device("cuda:0")
x = ops.as_nhwc(torch.randn(1, cin, H, W).to(dev))
out = ops.NHWC.empty(1, H, W, cout, dev)
if kind == "conv":
    w = torch.randn(cout, cin, k, k) * 0.05
    wp, ldw = packing.pack_conv(w.to(dev))
    sc, sh = torch.ones(cout, device=dev), torch.zeros(cout, device=dev)
    w_tc, bn = ops.tc_pack(wp, ldw, wp.shape[0], H * W, taps=k * k)
    fn = lambda: ops.conv2d_bn_act([x], wp, ldw, sc, sh, cout, k, k, 1, k // 2, True, out=out, flags=2, w_tc=w_tc, tc_bn=bn)
else:
    sp = {}
    spec._deform(sp, "d", cin, cout)
    sd = {kk: v.to(dev) for kk, v in spec.make_state_dict(sp, seed=3).items()}
    pk = packing.pack_dcn(sd, "d.conv", "d.actf.0")
    ws = ops.dcn_workspace(1, H, W, cin, cout, dev)
    fn = lambda: ops.dcnv2_forward(x, pk, relu=True, ws=ws, out=out, flags=2)
for _ in range(n):
    fn()
torch.cuda.synchronize()
print(case, "done")
