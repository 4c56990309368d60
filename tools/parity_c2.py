"""Full-size C2 frame (1088x608): per-tensor normalised error of both engines against the CPU oracle (GPU box; diagnostic).
   python tools/parity_c2.py [H W]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # tools/ sits next to deft_b200/
sys.path.insert(0, ROOT)
import torch
from oracle import restate, spec
from deft_b200.model.model import create_model
from deft_b200.opts import Opt

H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (608, 1088)
dev = torch.device("cuda:0")
heads = spec.MOT_HEADS
sd = spec.make_state_dict(spec.dlaseg_spec(heads))
x = spec.make_input((1, 3, H, W), tag="c2_mot_1088x608")
torch.set_num_threads(os.cpu_count() or 8)
with torch.no_grad():
    ref_out, ref_fm = restate.dlaseg_forward(sd, x, heads)
def nerr(a, b):
    a, b = a.double().cpu(), b.double()
    d = (a - b).abs()
    return float(d.max() / b.abs().max()), float(d.mean() / b.abs().mean()), float(b.abs().max())
res = {}
for name, tc in (("simt", False), ("tc3x", True)):
    opt = Opt(heads=dict(heads), tensor_core=tc)
    m = create_model("dla_34", dict(heads), {h: [256] for h in heads}, opt).eval()
    m.load_state_dict(sd, strict=True)
    m = m.to(dev)
    out, fm = m(x.to(dev))
    torch.cuda.synchronize()
    res[name] = ({h: out[0][h].clone() for h in heads}, [f.clone() for f in fm])
for i in range(13):
    print(f"fm{i:<2d} simt max/mean nerr %.2e %.2e | tc3x %.2e %.2e | tc3x vs simt %.2e | max|ref| %.2f" % (
        *nerr(res["simt"][1][i], ref_fm[i])[:2], *nerr(res["tc3x"][1][i], ref_fm[i])[:2], nerr(res["tc3x"][1][i], res["simt"][1][i].cpu())[0], nerr(res["simt"][1][i], ref_fm[i])[2]))
for h in heads:
    print(f"{h:12s} simt max/mean nerr %.2e %.2e | tc3x %.2e %.2e | tc3x vs simt %.2e | max|ref| %.2f" % (
        *nerr(res["simt"][0][h], ref_out[0][h])[:2], *nerr(res["tc3x"][0][h], ref_out[0][h])[:2], nerr(res["tc3x"][0][h], res["simt"][0][h].cpu())[0], nerr(res["simt"][0][h], ref_out[0][h])[2]))
