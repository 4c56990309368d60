"""Bring-up diagnostics for the tcgen05 engine (run on the GPU box; not a pytest file).

Each case runs in its own subprocess so a trapped kernel cannot poison the others; prints the
normalised max error of the tensor-core path (3xTF32 and 1xTF32) against the fp32 SIMT engine on the
same GPU, plus a small dump on mismatch.   python tools/tc_diag.py [case ...]
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # tools/ sits next to deft_b200/
sys.path.insert(0, ROOT)

CASES = {
    # name: (kind, cin, cout, k, stride, H, W, N)
    "c1x1_k32_n16": ("conv", 32, 16, 1, 1, 8, 16, 1),
    "c1x1_k32_n32": ("conv", 32, 32, 1, 1, 8, 16, 1),
    "c1x1_k32_n64": ("conv", 32, 64, 1, 1, 8, 16, 1),
    "c1x1_k32_n128": ("conv", 32, 128, 1, 1, 8, 16, 1),
    "c1x1_k128_n64": ("conv", 128, 64, 1, 1, 8, 16, 1),
    "c1x1_k256_n256": ("conv", 256, 256, 1, 1, 9, 13, 2),
    "c3x3_64_64": ("conv", 64, 64, 3, 1, 19, 23, 1),
    "c3x3_16_16": ("conv", 16, 16, 3, 1, 19, 23, 1),
    "c3x3_16_32_s2": ("conv", 16, 32, 3, 2, 20, 24, 2),
    "c3x3_128_128_s2": ("conv", 128, 128, 3, 2, 20, 24, 1),
    "c3x3_512_512": ("conv", 512, 512, 3, 1, 5, 7, 1),
    "root_448_128": ("root", 448, 128, 1, 1, 10, 14, 1),
    "dcn_64_64": ("dcn", 64, 64, 3, 1, 21, 30, 1),
    "dcn_512_256": ("dcn", 512, 256, 3, 1, 5, 7, 2),
    "dcn_128_128": ("dcn", 128, 128, 3, 1, 13, 10, 1),
    "heads_mot": ("heads", 64, 0, 3, 1, 21, 30, 2),
    "heads_nusc": ("heads_nusc", 64, 0, 3, 1, 21, 30, 1),
    "big_heads_mot": ("heads", 64, 0, 3, 1, 152, 272, 1),
    "c3x3_256_256_l4": ("conv", 256, 256, 3, 1, 38, 68, 1),
    "big_c3x3_64_64": ("conv", 64, 64, 3, 1, 152, 272, 1),
    "big_dcn_64_64": ("dcn", 64, 64, 3, 1, 152, 272, 1),
}


def run_case(name):
    import torch
    from deft_b200 import ops, packing
    from oracle import spec
    kind, cin, cout, k, stride, H, W, N = CASES[name]
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    res = {}
    for label, flags in (("simt", 0), ("tc3x", 2), ("tc1x", 2 | 4)):
        if kind in ("conv", "root"):
            w = (torch.rand(cout, cin, k, k) * 2 - 1) * (3.0 / (cin * k * k)) ** 0.5
            wp, ldw = packing.pack_conv(w.to(dev))
            sc = (torch.rand(cout) + 0.5).to(dev)
            sh = (torch.randn(cout) * 0.1).to(dev)
            if kind == "root":
                cs = [128, 128, 64, 128]
                xs = [spec.make_input((N, c, H, W), tag=f"r{i}") for i, c in enumerate(cs)]
            else:
                xs = [spec.make_input((N, cin, H, W), tag="x")]
            srcs = [ops.as_nhwc(x.to(dev)) for x in xs]
            OH = (H + 2 * (k // 2) - k) // stride + 1
            OW = (W + 2 * (k // 2) - k) // stride + 1
            resid = ops.as_nhwc(spec.make_input((N, cout, OH, OW), tag="res").to(dev))
            out = ops.conv2d_bn_act(srcs, wp, ldw, sc, sh, cout, k, k, stride, k // 2, True, residual=resid, flags=flags)
            torch.manual_seed(0)
        elif kind.startswith("heads"):
            heads = spec.NUSC_HEADS if kind == "heads_nusc" else spec.MOT_HEADS
            sd = spec.make_state_dict({kk: v for kk, v in spec.dlaseg_spec(heads).items() if kk.split(".")[0] in heads}, seed=21)
            hp = packing.pack_heads({kk: v.to(dev) for kk, v in sd.items()}, heads)
            feat = ops.as_nhwc(spec.make_input((N, 64, H, W), tag="hf").to(dev))
            outs = ops.heads_forward(feat, hp, flags=flags)
            torch.cuda.synchronize()
            res[label] = torch.cat([o.float().cpu() for o in outs], 1)
            continue
        else:
            sp = {}
            spec._deform(sp, "d", cin, cout)
            sd = {kk: v.to(dev) for kk, v in spec.make_state_dict(sp, seed=3).items()}
            pk = packing.pack_dcn(sd, "d.conv", "d.actf.0")
            x = ops.as_nhwc(spec.make_input((N, cin, H, W), tag="x").to(dev))
            out = ops.dcnv2_forward(x, pk, relu=True, flags=flags)
        torch.cuda.synchronize()
        res[label] = out.nchw_view().float().cpu().clone()
    ref = res["simt"]
    line = [name]
    for label in ("tc3x", "tc1x"):
        d = (res[label] - ref).abs()
        e = float(d.max() / ref.abs().max())
        line.append(f"{label}: nerr={e:.3e} mean={float(d.mean()):.3e}")
        if not e < (1e-4 if label == "tc3x" else 5e-3):
            bad = (d > 1e-3 * ref.abs().max()).nonzero()
            line.append(f"BAD n={bad.shape[0]}/{ref.numel()} first={bad[:6].tolist()}")
            n0, c0, y0, x0 = (bad[0].tolist() if bad.shape[0] else (0, 0, 0, 0))
            line.append("got=" + str(res[label][n0, :8, y0, x0].tolist()))
            line.append("ref=" + str(ref[n0, :8, y0, x0].tolist()))
            per_c = (d.amax(dim=(0, 2, 3)) > 1e-3 * ref.abs().max()).nonzero().view(-1).tolist()
            line.append("bad_channels=" + str(per_c[:40]))
            flat_bad_pix = (d.amax(dim=1) > 1e-3 * ref.abs().max()).view(-1).nonzero().view(-1).tolist()
            line.append("bad_pix(first40)=" + str(flat_bad_pix[:40]) + f" total={len(flat_bad_pix)}")
    print(" | ".join(line), flush=True)


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        run_case(sys.argv[2])
        return
    names = sys.argv[1:] or list(CASES)
    for n in names:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", n], capture_output=True, text=True, timeout=120)
            out = (r.stdout.strip() or "") + ((" STDERR: " + r.stderr.strip()[-600:]) if r.returncode != 0 else "")
            print(out if out else f"{n}: no output rc={r.returncode}", flush=True)
        except subprocess.TimeoutExpired:
            print(f"{n}: TIMEOUT", flush=True)


if __name__ == "__main__":
    main()
