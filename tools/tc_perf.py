"""Perf bring-up for the tcgen05 engine (GPU box; not a pytest file): times single layers with CUDA events.
   DEFT_TC_DBG=<bits> python tools/tc_perf.py   (bits: 1 no B copy, 2 no A loads, 4 no A stores, 8 no MMA, 16 no epilogue, 32 arm the progress board for stall reports)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # tools/ sits next to deft_b200/
sys.path.insert(0, ROOT)
import torch
from deft_b200 import ops, packing
from oracle import spec

dev = torch.device("cuda:0")
CASES = [("c3x3 64->64 @152x272", "conv", 64, 64, 3, 152, 272), ("c1x1 64->64 @152x272", "conv", 64, 64, 1, 152, 272),
         ("c3x3 128->128 @76x136", "conv", 128, 128, 3, 76, 136), ("c3x3 256->256 @38x68", "conv", 256, 256, 3, 38, 68),
         ("dcn 64->64 @152x272", "dcn", 64, 64, 3, 152, 272), ("dcn 128->64 @76x136", "dcn", 128, 64, 3, 76, 136),
         ("dcn 128->128 @76x136", "dcn", 128, 128, 3, 76, 136), ("dcn 256->256 @38x68", "dcn", 256, 256, 3, 38, 68),
         ("dcn 256->128 @38x68", "dcn", 256, 128, 3, 38, 68),
         ("heads mot @152x272", "heads", 64, 0, 3, 152, 272)]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for name, kind, cin, cout, k, H, W in CASES:
    if kind == "conv":
        w = torch.randn(cout, cin, k, k) * 0.05
        wp, ldw = packing.pack_conv(w.to(dev))
        sc, sh = torch.ones(cout, device=dev), torch.zeros(cout, device=dev)
        x = ops.as_nhwc(torch.randn(1, cin, H, W).to(dev))
        out = ops.NHWC.empty(1, H, W, cout, dev)
        M = H * W
        w_tc, bn = ops.tc_pack(wp, ldw, wp.shape[0], M, taps=k * k)
        ws = ops.tc_scratch(ldw, wp.shape[0], M, 0, dev)
        fn = lambda: ops.conv2d_bn_act([x], wp, ldw, sc, sh, cout, k, k, 1, k // 2, True, out=out, flags=2, w_tc=w_tc, tc_bn=bn, ws=ws)
    elif kind == "dcn":
        sp = {}
        spec._deform(sp, "d", cin, cout)
        sd = {kk: v.to(dev) for kk, v in spec.make_state_dict(sp, seed=3).items()}
        pk = packing.pack_dcn(sd, "d.conv", "d.actf.0")
        x = ops.as_nhwc(torch.randn(1, cin, H, W).to(dev))
        out = ops.NHWC.empty(1, H, W, cout, dev)
        ws = ops.dcn_workspace(1, H, W, cin, cout, dev)
        fn = lambda: ops.dcnv2_forward(x, pk, relu=True, ws=ws, out=out, flags=2)
    else:
        heads = spec.MOT_HEADS
        sd = spec.make_state_dict({kk: v for kk, v in spec.dlaseg_spec(heads).items() if kk.split(".")[0] in heads}, seed=21)
        hp = packing.pack_heads({kk: v.to(dev) for kk, v in sd.items()}, heads)
        feat = ops.as_nhwc(torch.randn(1, 64, H, W).to(dev))
        outs = [torch.empty((1, c, H, W), device=dev) for c in hp["head_c"]]
        wsb = torch.empty(256, dtype=torch.uint8, device=dev)
        fn = lambda: ops.heads_forward(feat, hp, outs=outs, ws=wsb, flags=2)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    print(f"dbg={os.environ.get('DEFT_TC_DBG','0'):>3s} stages={os.environ.get('DEFT_TC_STAGES','-'):>2s} {name:26s} median {ts[5]:8.1f} us  min {ts[0]:8.1f} us", flush=True)
