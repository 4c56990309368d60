"""Protocol timeline of the tcgen05 TS kernel (GPU box; perf bring-up): CTA 0 stamps clock64() at each protocol point.
   python tools/tc_trace.py [conv64|conv128|conv1x1|dcn64|dcn128]   -> per-role tables + a per-k-block summary."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # tools/ sits next to deft_b200/
sys.path.insert(0, ROOT)
import torch
from deft_b200 import ops, packing
from oracle import spec

CASES = {"conv32": ("conv", 64, 32, 3, 152, 272), "conv64": ("conv", 64, 64, 3, 152, 272), "conv128": ("conv", 128, 128, 3, 76, 136), "conv1x1": ("conv", 128, 64, 1, 152, 272),
         "conv16": ("conv", 16, 16, 3, 608, 1088), "dcn64": ("dcn", 64, 64, 3, 152, 272), "dcn128": ("dcn", 128, 128, 3, 76, 136)}
case = sys.argv[1] if len(sys.argv) > 1 else "conv64"
kind, cin, cout, k, H, W = CASES[case]
dev = torch.device("cuda:0")
trace = torch.zeros(4 * 96 * 4, dtype=torch.int64, device=dev)
x = ops.as_nhwc(torch.randn(1, cin, H, W).to(dev))
out = ops.NHWC.empty(1, H, W, cout, dev)
if kind == "conv":
    w = torch.randn(cout, cin, k, k) * 0.05
    wp, ldw = packing.pack_conv(w.to(dev))
    sc, sh = torch.ones(cout, device=dev), torch.zeros(cout, device=dev)
    M = H * W
    w_tc, bn = ops.conv_tc_pack(wp, ldw, cin, k, k, M)
    fn = lambda: ops.conv2d_bn_act([x], wp, ldw, sc, sh, cout, k, k, 1, k // 2, True, out=out, flags=2, w_tc=w_tc, tc_bn=bn)
else:
    sp = {}
    spec._deform(sp, "d", cin, cout)
    sd = {kk: v.to(dev) for kk, v in spec.make_state_dict(sp, seed=3).items()}
    pk = packing.pack_dcn(sd, "d.conv", "d.actf.0")
    ws = ops.dcn_workspace(1, H, W, cin, cout, dev)
    fn = lambda: ops.dcnv2_forward(x, pk, relu=True, ws=ws, out=out, flags=2)
for _ in range(3):
    fn()
torch.cuda.synchronize()
from deft_b200 import _lib
_lib.lib().deft_debug_set_trace(trace.data_ptr())
fn()
torch.cuda.synchronize()
_lib.lib().deft_debug_set_trace(None)
t = trace.cpu().view(4, 96, 4)
t0 = int(t[t > 0].min())
names = ["MMA(issuer 0): loop_top  a_full_ok  b_full_ok  committed", "PROD(group leader): top  filled  a_empty_ok  arrived", "BCOPY: top  b_empty_ok",
         "EPI: top  acc_full_ok  done"]
print(f"== {case}: {kind} {cin}->{cout} k{k} @{H}x{W}  (for dcn: the LAST traced launch is the gather-GEMM; the offset conv's stamps are overwritten)")
for role in range(4):
    print(names[role])
    n = 60 if role < 3 else 24
    for i in range(n):
        row = t[role, i]
        if int(row.max()) == 0:
            continue
        print(f"  {i:3d} " + " ".join(f"{(int(v) - t0) if int(v) > 0 else -1:8d}" for v in row))
# summary: producer phases (cycles) averaged over the traced k-blocks, issuer phases over its own k-blocks
def avg(role, a, b, lo=4, hi=60):
    d = [int(t[role, i, b] - t[role, i, a]) for i in range(lo, hi) if int(t[role, i, a]) > 0 and int(t[role, i, b]) > 0]
    return sum(d) / max(len(d), 1), len(d)
print("summary (cycles, k-blocks 4..59):")
print("  producer fill (top->filled)      %.0f  (n=%d)" % avg(1, 0, 1))
print("  producer wait a_empty            %.0f" % avg(1, 1, 2)[0])
print("  producer split+tcgen05.st+arrive %.0f" % avg(1, 2, 3)[0])
print("  issuer wait a_full               %.0f  (n=%d)" % avg(0, 0, 1))
print("  issuer wait b_full               %.0f" % avg(0, 1, 2)[0])
print("  issuer 12 MMAs + commits         %.0f" % avg(0, 2, 3)[0])
own = [i for i in range(4, 60) if int(t[0, i, 0]) > 0]
if len(own) > 2:
    print("  issuer 0 period per own k-block  %.0f  -> per k-block %.0f" % ((int(t[0, own[-1], 0]) - int(t[0, own[0], 0])) / (len(own) - 1),
                                                                            (int(t[0, own[-1], 0]) - int(t[0, own[0], 0])) / (own[-1] - own[0])))
