"""Stress of the affinity GEMM chain at the C4 problem size (F=4, rows=250, K=300 -> 300000 pair rows), many iterations, several
streams concurrently: the configuration in which a tc_ts_kernel<TsConvAP,64> stalled once (round 2).  python tools/stress_affinity.py [iters]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # tools/ sits next to deft_b200/
sys.path.insert(0, ROOT)
import torch
from deft_b200 import ops, packing
from oracle import spec

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 400
dev = torch.device("cuda:0")
sd = {k[4:]: v.to(dev) for k, v in spec.make_state_dict(spec.dlaseg_spec(spec.NUSC_HEADS, dataset="nuscenes")).items() if k.startswith("AFE.")}
pk = packing.pack_affinity(sd)
F, P, rows, K = 4, 300, 250, 300
streams = [torch.cuda.Stream() for _ in range(3)]
state = []
for st in streams:
    with torch.cuda.stream(st):
        au = torch.randn(F, P, 512, device=dev)
        bv = torch.randn(K, 512, device=dev)
        n = torch.full((F,), rows, dtype=torch.int32, device=dev)
        out = torch.empty(F, P, P + 1, device=dev)
        ws = torch.empty(ops.lib().deft_afe_pairs_workspace_bytes(F, rows, K), dtype=torch.uint8, device=dev)
        state.append((au, bv, n, out, ws))
torch.cuda.synchronize()
for it in range(iters):
    for st, (au, bv, n, out, ws) in zip(streams, state):
        with torch.cuda.stream(st):
            ops.afe_affinity_cached(au, n, bv, K, P, pk, rows=rows, out=out, ws=ws, flags=2)
    if it % 50 == 49:
        torch.cuda.synchronize()
        print("iter", it + 1, "ok", float(out.sum()), flush=True)
torch.cuda.synchronize()
print("stress_affinity done")
