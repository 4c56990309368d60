"""GPU parity tests, per operator: every C-ABI entry point against the CPU oracle (oracle/restate.py)
and the committed reference goldens, on the same seeded inputs.  Tolerance: 1e-3 normalised max error
(BASELINE.json) -- the fp32 SIMT anchor is held to 2e-5; bit-exact for indices/scores/boxes.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN, nerr
from oracle import restate, spec

pytestmark = pytest.mark.gpu

TOL_FP32 = 2e-5
TOL = 1e-3


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from deft_b200 import _lib
    assert _lib.lib().deft_device_ok(), _lib.lib().deft_last_error()
    return torch.device("cuda:0")


def rnd(shape, tag, scale=1.0):
    return spec.make_input(shape, tag=tag) * scale


ENGINES = [pytest.param(0, id="simt"), pytest.param(2, id="tc3x")]


def tol_for(flags):
    return TOL_FP32 if flags == 0 else 1e-4


def test_layout_roundtrip(dev):
    from deft_b200 import ops
    x = rnd((2, 37, 13, 29), "layout").to(dev)
    a = ops.as_nhwc(x)
    assert torch.equal(a.t.permute(0, 3, 1, 2).cpu(), x.cpu())
    assert torch.equal(ops.to_nchw(a).cpu(), x.cpu())
    xcl = x.contiguous(memory_format=torch.channels_last)
    assert xcl.data_ptr() != x.data_ptr()


@pytest.mark.parametrize("flags", ENGINES)
@pytest.mark.parametrize("cin,cout,k,stride,hw", [(16, 16, 3, 1, (20, 28)), (16, 32, 3, 2, (20, 28)), (32, 64, 3, 2, (18, 22)),
                                                  (64, 64, 3, 1, (17, 19)), (128, 128, 3, 1, (9, 12)), (64, 128, 1, 1, (11, 13)),
                                                  (256, 256, 3, 1, (6, 8)), (512, 512, 3, 1, (3, 4)),
                                                  # 16x8-patch (TMA halo) producer: exact / ragged tilings, split-K, narrow N
                                                  (64, 64, 3, 1, (30, 62)), (128, 64, 3, 1, (16, 32)), (256, 128, 3, 1, (30, 62)),
                                                  (64, 32, 3, 1, (32, 48)),
                                                  # 16-channel halo patch (TsPatch16AP: DLA level0): exact / ragged tilings, N = 16 and 32
                                                  (16, 16, 3, 1, (32, 48)), (16, 16, 3, 1, (30, 62)), (16, 32, 3, 1, (24, 64)),
                                                  (16, 16, 3, 1, (160, 256))])      # 640 patches: 4-5 tiles per CTA (the patch ring wraps)
def test_conv_bn_act(dev, cin, cout, k, stride, hw, flags):
    from deft_b200 import ops, packing
    sp = {}
    sp["c.weight"] = ("conv", (cout, cin, k, k))
    spec._bn(sp, "b", cout)
    sd = spec.make_state_dict(sp, seed=cin * 7 + cout)
    x = rnd((2, cin, *hw), f"conv{cin}{cout}{k}")
    res = rnd((2, cout, (hw[0] + 2 * (k // 2) - k) // stride + 1, (hw[1] + 2 * (k // 2) - k) // stride + 1), "res")
    ref = F.relu(restate.bn_eval(sd, "b", F.conv2d(x, sd["c.weight"], stride=stride, padding=k // 2)) + res)
    wp, ldw = packing.pack_conv(sd["c.weight"].to(dev))
    sc, sh = packing.fold_bn({k_: v.to(dev) for k_, v in sd.items()}, "b")
    out = ops.conv2d_bn_act([ops.as_nhwc(x.to(dev))], wp, ldw, sc, sh, cout, k, k, stride, k // 2, True,
                            residual=ops.as_nhwc(res.to(dev)), flags=flags)
    assert nerr(out.nchw_view().cpu(), ref) < tol_for(flags)


@pytest.mark.parametrize("cout,stride,hw", [(16, 1, (37, 70)), (16, 1, (64, 128)), (32, 2, (38, 70)), (32, 2, (64, 96))])
def test_conv_c16_direct_kernel(dev, cout, stride, hw):
    """The specialised direct kernel for DLA level0 (16->16) / level1 (16->32, stride 2): no residual, odd sizes,
    batch 2; also cross-checked against the generic fp32 engine (DEFT_FLAG_NO_DIRECT = 8)."""
    from deft_b200 import ops, packing
    sp = {"c.weight": ("conv", (cout, 16, 3, 3))}
    spec._bn(sp, "b", cout)
    sd = spec.make_state_dict(sp, seed=cout + stride)
    x = rnd((2, 16, *hw), f"c16{cout}")
    ref = F.relu(restate.bn_eval(sd, "b", F.conv2d(x, sd["c.weight"], stride=stride, padding=1)))
    wp, ldw = packing.pack_conv(sd["c.weight"].to(dev))
    sc, sh = packing.fold_bn({k_: v.to(dev) for k_, v in sd.items()}, "b")
    out = ops.conv2d_bn_act([ops.as_nhwc(x.to(dev))], wp, ldw, sc, sh, cout, 3, 3, stride, 1, True)
    gen = ops.conv2d_bn_act([ops.as_nhwc(x.to(dev))], wp, ldw, sc, sh, cout, 3, 3, stride, 1, True, flags=8)
    assert nerr(out.nchw_view().cpu(), ref) < TOL_FP32
    assert nerr(gen.nchw_view().cpu(), ref) < TOL_FP32


@pytest.mark.parametrize("flags", ENGINES)
def test_conv_concat_root(dev, flags):
    """Root: 1x1 conv over a channel concat of 4 sources (dla.py:199-207) without materialising the cat."""
    from deft_b200 import ops, packing
    cs = [128, 128, 64, 128]
    sp = {"c.weight": ("conv", (128, sum(cs), 1, 1))}
    spec._bn(sp, "b", 128)
    sd = spec.make_state_dict(sp, seed=5)
    xs = [rnd((1, c, 10, 14), f"root{i}") for i, c in enumerate(cs)]
    ref = F.relu(restate.bn_eval(sd, "b", F.conv2d(torch.cat(xs, 1), sd["c.weight"])))
    wp, ldw = packing.pack_conv(sd["c.weight"].to(dev))
    sc, sh = packing.fold_bn({k: v.to(dev) for k, v in sd.items()}, "b")
    out = ops.conv2d_bn_act([ops.as_nhwc(x.to(dev)) for x in xs], wp, ldw, sc, sh, 128, 1, 1, 1, 0, True, flags=flags)
    assert nerr(out.nchw_view().cpu(), ref) < tol_for(flags)


def test_stem_maxpool_upsample(dev):
    from deft_b200 import ops, packing
    sp = {"c.weight": ("conv", (16, 3, 7, 7))}
    spec._bn(sp, "b", 16)
    sd = spec.make_state_dict(sp, seed=9)
    x = rnd((2, 3, 37, 50), "stem")
    ref = F.relu(restate.bn_eval(sd, "b", F.conv2d(x, sd["c.weight"], padding=3)))
    wp, _ = packing.pack_conv(sd["c.weight"].to(dev), 16)
    sc, sh = packing.fold_bn({k: v.to(dev) for k, v in sd.items()}, "b")
    out = ops.stem_conv7x7(x.to(dev), wp, sc, sh)
    assert nerr(out.nchw_view().cpu(), ref) < TOL_FP32
    y = rnd((2, 32, 12, 18), "pool")
    assert torch.equal(ops.maxpool2x2(ops.as_nhwc(y.to(dev))).nchw_view().cpu(), F.max_pool2d(y, 2, 2))
    for f in (2, 4):
        spu = {"u.weight": ("up", (64, 1, 2 * f, 2 * f))}
        w = spec.make_state_dict(spu, seed=f)["u.weight"]
        a = rnd((2, 64, 5, 7), f"up{f}")
        skip = rnd((2, 64, 5 * f, 7 * f), f"skip{f}")
        ref = F.conv_transpose2d(a, w, stride=f, padding=f // 2, groups=64) + skip
        wpu, ff = packing.pack_deconv(w.to(dev))
        assert ff == f
        out = ops.upsample_add(ops.as_nhwc(a.to(dev)), wpu, f, skip=ops.as_nhwc(skip.to(dev)))
        assert nerr(out.nchw_view().cpu(), ref) < TOL_FP32


@pytest.mark.parametrize("flags", ENGINES)
def test_dcn_layers_vs_reference_golden(dev, flags):
    """DeformConv (DCN+BN+ReLU) against outputs of the reference's own module (tests/golden/dcn_layers.pt)."""
    from deft_b200 import ops, packing
    cases = torch.load(os.path.join(GOLDEN, "dcn_layers.pt"), weights_only=False)
    for c in cases:
        sp = {}
        spec._deform(sp, "d", c["chi"], c["cho"])
        sd = {k: v.to(dev) for k, v in spec.make_state_dict(sp, seed=c["seed"]).items()}
        x = spec.make_input((1, c["chi"], c["h"], c["w"]), tag=f"dcn{c['chi']}x{c['cho']}").to(dev)
        pk = packing.pack_dcn(sd, "d.conv", "d.actf.0")
        y = ops.dcnv2_forward(ops.as_nhwc(x), pk, relu=True, flags=flags).nchw_view().cpu()
        assert nerr(y, c["y"]) < tol_for(flags), (c["chi"], c["cho"])


@pytest.mark.parametrize("flags", ENGINES)
@pytest.mark.parametrize("chi,cho,h,w", [(512, 256, 4, 6), (256, 256, 7, 9), (256, 128, 8, 6), (128, 128, 13, 10),
                                         (128, 64, 14, 9), (64, 64, 21, 30), (256, 64, 8, 9)])
def test_dcn_all_dla_shapes_vs_oracle(dev, chi, cho, h, w, flags):
    """Every (Cin,Cout) pair DLA-34's 16 DCN layers use, batch 2, odd sizes, with
    offsets large enough to leave the image; bare op (no BN/ReLU) through the drop-in DCN module."""
    from deft_b200.dcn_v2 import DCN
    sp = {}
    spec._deform(sp, "d", chi, cho)
    sd = spec.make_state_dict(sp, seed=chi + cho + h)
    sd["d.conv.conv_offset_mask.bias"] = sd["d.conv.conv_offset_mask.bias"] * 3.0
    x = rnd((2, chi, h, w), f"dcnall{chi}{cho}")
    with torch.no_grad():
        ref = restate.dcn_forward(sd, "d.conv", x, explicit=False)
    mod = DCN(chi, cho)
    mod.load_state_dict({k[len("d.conv."):]: v for k, v in sd.items() if k.startswith("d.conv.")})
    mod = mod.to(dev)
    with torch.no_grad():
        y = mod(x.to(dev), flags=flags)
    assert nerr(y.cpu(), ref) < tol_for(flags)


def test_dcn_zero_offset_identity(dev):
    """Upstream init (conv_offset_mask = 0) => DCN(x) == 0.5*conv2d(x,W,pad=1)+b."""
    from deft_b200.dcn_v2 import DCN
    mod = DCN(64, 32).to(dev)
    with torch.no_grad():
        mod.bias.normal_()
        mod._invalidate()
        x = rnd((1, 64, 19, 23), "dcnid").to(dev)
        y = mod(x).cpu()
        ref = 0.5 * F.conv2d(x.cpu(), mod.weight.cpu(), None, padding=1) + mod.bias.cpu().view(1, -1, 1, 1)
    assert nerr(y, ref) < TOL_FP32


@pytest.mark.parametrize("flags", ENGINES)
@pytest.mark.parametrize("heads", [spec.MOT_HEADS, spec.NUSC_HEADS], ids=["mot", "nuscenes"])
def test_heads(dev, heads, flags):
    from deft_b200 import ops, packing
    sp = spec.dlaseg_spec(heads)
    sd = spec.make_state_dict({k: v for k, v in sp.items() if k.split(".")[0] in heads}, seed=21)
    feat = rnd((2, 64, 14, 17), "headfeat")
    hp = packing.pack_heads({k: v.to(dev) for k, v in sd.items()}, heads)
    outs = ops.heads_forward(ops.as_nhwc(feat.to(dev)), hp, flags=flags)
    for h, o in zip(hp["names"], outs):
        assert nerr(o.cpu(), restate.head_forward(sd, h, feat)) < tol_for(flags), h


def _check_decode(dec_gpu, dec_ref, exact_order=True):
    assert torch.equal(dec_gpu["scores"].cpu(), dec_ref["scores"])
    keys = ["clses", "xs", "ys", "cts", "bboxes", "tracking", "bboxes_amodal", "dep", "rot", "dim", "amodel_offset"]
    for k in keys:
        if k in dec_ref:
            assert k in dec_gpu, k
            assert torch.equal(dec_gpu[k].cpu(), dec_ref[k]), k


@pytest.mark.parametrize("name", ["frame_mot_96x128", "frame_kitti_64x160", "frame_nusc_64x96"])
def test_decode_bit_exact_on_reference_heads(dev, name):
    """generic_decode on the reference's own heads (shared tensor) -> bit-exact indices, scores, boxes."""
    from deft_b200.model.decode import generic_decode
    g = torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)
    sig = restate.sigmoid_output(g["heads"])
    out = generic_decode({k: v.to(dev) for k, v in sig.items()}, K=g["meta"]["K"], opt=None)
    _check_decode(out, g["decode"])
    ref = restate.generic_decode(sig, K=g["meta"]["K"])
    _check_decode(out, ref)


def test_decode_plateau_and_sizes(dev):
    from deft_b200.model.decode import generic_decode
    g = torch.load(os.path.join(GOLDEN, "decode_plateau.pt"), weights_only=False)
    out = generic_decode({k: v.to(dev) for k, v in g["in"].items()}, K=g["K"], opt=None)
    ref = restate.generic_decode(g["in"], K=g["K"])          # lowest-flat-index tie rule
    _check_decode(out, ref)
    assert torch.equal(out["scores"].cpu(), g["decode"]["scores"])
    # larger random case, K=100, 3 classes, odd sizes; plus K == H*W edge
    for (B, C, H, W, K) in ((2, 3, 37, 53, 100), (1, 1, 152, 272, 100), (1, 2, 4, 5, 20), (1, 10, 58, 102, 300)):
        hm = torch.sigmoid(rnd((B, C, H, W), f"hm{H}"))
        o = {"hm": hm, "reg": rnd((B, 2, H, W), "reg").sigmoid(), "wh": rnd((B, 2, H, W), "wh"),
             "tracking": rnd((B, 2, H, W), "trk"), "ltrb_amodal": rnd((B, 4, H, W), "ltrb")}
        out = generic_decode({k: v.to(dev) for k, v in o.items()}, K=K, opt=None)
        _check_decode(out, restate.generic_decode(o, K=K))


def test_sigmoid_output(dev):
    from deft_b200.model.decode import sigmoid_output
    hm, dep = rnd((1, 3, 9, 11), "sghm") * 3, rnd((1, 1, 9, 11), "sgdep")
    ref = restate.sigmoid_output({"hm": hm, "dep": dep})
    out = sigmoid_output({"hm": hm.to(dev), "dep": dep.to(dev)})
    assert nerr(out["hm"].cpu(), ref["hm"]) < 1e-6
    assert nerr(out["dep"].cpu(), ref["dep"]) < 1e-5


@pytest.mark.parametrize("dataset,heads", [("mot", spec.MOT_HEADS), ("nuscenes", spec.NUSC_HEADS)])
def test_afe_extract_vs_oracle(dev, dataset, heads):
    from deft_b200.model.networks.AFE import AFE_module
    from deft_b200.opts import Opt
    sp = {k: v for k, v in spec.dlaseg_spec(heads, dataset=dataset).items() if k.startswith("AFE.")}
    sd = spec.make_state_dict(sp, seed=31)
    H, W = 64, 96
    chans = spec.SELECTOR_IN
    scale = [1, 2, 4, 8, 16, 32, 4, 8, 16, 32, 4, 4, 4]
    fms = [rnd((1, c, H // s, W // s), f"fm{i}") for i, (c, s) in enumerate(zip(chans, scale))]
    N = 23
    centres = spec.make_uniform((1, N, 1, 1, 2), -1.1, 1.1, tag="afe_c")
    centres[0, 0, 0, 0] = torch.tensor([-1.0, -1.0]); centres[0, 1, 0, 0] = torch.tensor([1.0, 1.0])
    centres[0, 2, 0, 0] = torch.tensor([0.0, 0.0])
    with torch.no_grad():
        ref = restate.afe_extract(sd, fms, centres)
    mod = AFE_module(Opt(dataset=dataset))
    mod.load_state_dict({k[4:]: v for k, v in sd.items()})
    mod = mod.to(dev)
    emb = mod.forward_feature_extracter([f.to(dev) for f in fms], centres.to(dev))
    assert emb.shape == ref.shape
    assert nerr(emb.cpu(), ref) < TOL_FP32 * 5


@pytest.mark.parametrize("flags", ENGINES)
@pytest.mark.parametrize("dataset,heads,P,n,m", [("mot", spec.MOT_HEADS, 100, 7, 10), ("mot", spec.MOT_HEADS, 100, 100, 100),
                                                 ("mot", spec.MOT_HEADS, 100, 1, 3), ("nuscenes", spec.NUSC_HEADS, 30, 30, 9)])
def test_affinity_vs_oracle(dev, dataset, heads, P, n, m, flags):
    from deft_b200.model.networks.AFE import AFE_module
    from deft_b200.opts import Opt
    sp = {k: v for k, v in spec.dlaseg_spec(heads, dataset=dataset).items() if k.startswith("AFE.")}
    sd = spec.make_state_dict(sp, seed=33)
    cemb = sum(spec.selector_out(dataset))
    xp = rnd((1, n, cemb), "affp").abs()
    xn = rnd((1, m, cemb), "affn").abs()
    mod = AFE_module(Opt(dataset=dataset, max_object=P, tensor_core=bool(flags)))
    mod.load_state_dict({k[4:]: v for k, v in sd.items()})
    mod = mod.to(dev)
    for fill in (False, True):
        with torch.no_grad():
            ref = restate.afe_affinity(sd, xp, xn, P, fill)
        out = mod.forward_stacker_features(xp.to(dev), xn.to(dev), fill)
        assert out.shape == ref.shape and out.dtype == np.float32
        assert nerr(out, ref) < (1e-5 if flags == 0 else 1e-4)


def test_affinity_batched_frames_and_golden(dev):
    """F memory frames with ragged counts in one call == F separate reference calls (tracker.py:76-90)."""
    from deft_b200.model.networks.AFE import AFE_module
    from deft_b200.opts import Opt
    g = torch.load(os.path.join(GOLDEN, "frame_mot_96x128.pt"), weights_only=False)
    sd = spec.make_state_dict({k: v for k, v in spec.dlaseg_spec(spec.MOT_HEADS).items() if k.startswith("AFE.")})
    mod = AFE_module(Opt(tensor_core=False))
    mod.load_state_dict({k[4:]: v for k, v in sd.items()})
    mod = mod.to(dev)
    assert nerr(mod.forward_stacker_features(g["emb_prev"].to(dev), g["emb"].to(dev), False), g["aff"]) < 1e-5
    assert nerr(mod.forward_stacker_features(g["emb_prev"].to(dev), g["emb"].to(dev), True), g["aff_fill"]) < 1e-5
    assert nerr(mod.forward_stacker_features(g["emb"].to(dev), g["emb"].to(dev), False), g["aff_full"]) < 1e-5
    counts = [5, 0, 100, 37]
    P, m = 100, 42
    mem = rnd((4, P, 416), "mem").abs()
    cur = rnd((m, 416), "cur").abs()
    out = mod.affinity_batched(mem.to(dev), counts, cur.to(dev)).cpu()
    for f, n in enumerate(counts):
        if n == 0:
            continue
        ref = restate.afe_affinity(sd, mem[f:f + 1, :n], cur.unsqueeze(0), P, False)
        assert nerr(out[f, :n, : m + 1], ref) < 1e-5, f


def test_lstm_step_vs_golden_and_oracle(dev):
    from deft_b200.kalman_filter_lstm import DecoderRNN, KalmanFilterLSTM
    from deft_b200.opts import Opt
    g = torch.load(os.path.join(GOLDEN, "lstm_step.pt"), weights_only=False)
    for dataset, fin, fut in (("mot", 11, 5), ("nuscenes", 18, 4)):
        sd = spec.make_state_dict(spec.lstm_spec(dataset), seed=55)
        model = DecoderRNN(128, dataset)
        model.load_state_dict(sd)
        model = model.to(dev)
        kf = KalmanFilterLSTM(Opt(dataset=dataset), model)
        T = g[dataset]["T"]
        h0 = (spec.make_input((T, 128), tag="h0") * 0.1).to(dev)
        c0 = (spec.make_input((T, 128), tag="c0") * 0.1).to(dev)
        x = spec.make_input((T, fin), tag="lstm_x").to(dev)
        hn, cn, pred = kf.predict_batch(h0, c0, x)
        assert nerr(hn.cpu(), g[dataset]["hn"]) < 1e-5
        assert nerr(cn.cpu(), g[dataset]["cn"]) < 1e-5
        assert nerr(pred.cpu(), g[dataset]["pred"]) < 1e-5
        # the reference's one-track signature
        h1, c1, d = kf.predict(h0[2].view(1, 1, 128), c0[2].view(1, 1, 128), x[2].view(1, 1, fin))
        assert sorted(d.keys()) == list(range(1, fut + 1)) and d[1].shape == (4,)
        assert nerr(torch.from_numpy(np.stack([d[i] for i in range(1, fut + 1)])), g[dataset]["pred"][2]) < 1e-5
        assert nerr(h1.view(128).cpu(), g[dataset]["hn"][2]) < 1e-5
        # big batch vs oracle (T=500 tracks)
        T2 = 500
        hb, cb, xb = rnd((T2, 128), "hb", 0.1), rnd((T2, 128), "cb", 0.1), rnd((T2, fin), "xb")
        rh, rc, rp = restate.lstm_predict(sd, hb, cb, xb, fut)
        gh, gc, gp = kf.predict_batch(hb.to(dev), cb.to(dev), xb.to(dev))
        assert nerr(gh.cpu(), rh) < 1e-5 and nerr(gc.cpu(), rc) < 1e-5 and nerr(gp.cpu(), rp) < 1e-5
