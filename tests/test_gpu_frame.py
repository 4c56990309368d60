"""GPU parity tests, whole frame: the drop-in DLASeg / AFE / decode / LSTM path against (1) goldens produced
by the REAL reference modules (tests/golden/frame_*.pt), (2) the CPU oracle at BASELINE config C1
(512x512, 10x10 affinity), (3) size-independent properties at the full C2 size (1088x608)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, nerr, oracle_threads
from oracle import restate, spec

pytestmark = pytest.mark.gpu
TOL = 1e-3          # BASELINE.json: 1e-3 relative fp32
TOL_ANCHOR = 5e-5   # the fp32 SIMT engine


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    return torch.device("cuda:0")


def build(heads, dataset, dev, max_object=100, **kw):
    from deft_b200.model.model import create_model
    from deft_b200.opts import Opt
    opt = Opt(dataset=dataset, max_object=max_object, heads=dict(heads), **kw)
    m = create_model("dla_34", dict(heads), {h: [256] for h in heads}, opt).eval()
    sd = spec.make_state_dict(spec.dlaseg_spec(heads, dataset=dataset))
    missing = m.load_state_dict(sd, strict=True)
    return m.to(dev), sd, opt


FRAMES = [("frame_mot_96x128", spec.MOT_HEADS, "mot"), ("frame_kitti_64x160", spec.KITTI_HEADS, "kitti"),
          ("frame_nusc_64x96", spec.NUSC_HEADS, "nuscenes")]


@pytest.mark.parametrize("tensor_core", [False, True], ids=["simt", "tc3x"])
@pytest.mark.parametrize("name,heads,dataset", FRAMES)
def test_frame_vs_reference_golden(dev, name, heads, dataset, tensor_core):
    g = torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)
    mt = g["meta"]
    m, sd, opt = build(heads, dataset, dev, max_object=mt["max_object"], tensor_core=tensor_core)
    tol = TOL if tensor_core else TOL_ANCHOR
    x = spec.make_input((1, 3, mt["H"], mt["W"]), tag="image").to(dev)
    out, fm = m(x)
    assert isinstance(out, list) and len(out) == 1 and list(out[0].keys()) == list(heads.keys())
    errs = {}
    for h in heads:
        assert out[0][h].shape == g["heads"][h].shape
        errs[h] = nerr(out[0][h].cpu(), g["heads"][h])
    assert len(fm) == 13
    for i, (f, d) in enumerate(zip(fm, g["fm"])):
        assert list(f.shape) == d["shape"], i
        fc = f.contiguous().cpu().view(-1)
        errs[f"fm{i}"] = nerr(fc[d["pos"]], d["val"])
        assert abs(float(fc.double().abs().sum()) - d["abssum"]) / d["abssum"] < 1e-3
    centres = spec.make_uniform((1, mt["N"], 1, 1, 2), -1.05, 1.05, tag="centres").to(dev)
    centres_prev = spec.make_uniform((1, mt["n_prev"], 1, 1, 2), -1.0, 1.0, tag="centres_prev").to(dev)
    emb = m.AFE.forward_feature_extracter(fm, centres)
    emb_prev = m.AFE.forward_feature_extracter(fm, centres_prev)
    errs["emb"] = nerr(emb.cpu(), g["emb"])
    errs["emb_prev"] = nerr(emb_prev.cpu(), g["emb_prev"])
    errs["aff"] = nerr(m.AFE.forward_stacker_features(emb_prev, emb, False), g["aff"])
    errs["aff_fill"] = nerr(m.AFE.forward_stacker_features(emb_prev, emb, True), g["aff_fill"])
    print(name, "tc" if tensor_core else "simt", {k: f"{v:.2e}" for k, v in errs.items()})
    bad = {k: v for k, v in errs.items() if not v < tol}
    assert not bad, bad
    # decode on OUR heads: same detections as the reference's decode on ITS heads wherever the score gap
    # to the next candidate exceeds the network tolerance (hard part 4)
    from deft_b200.model.decode import generic_decode, sigmoid_output
    dec = generic_decode(sigmoid_output({k: v.clone() for k, v in out[0].items()}, opt), K=mt["K"], opt=opt)
    ref = g["decode"]
    assert nerr(dec["scores"].cpu(), ref["scores"]) < 5 * tol
    gaps = (ref["scores"][0, :-1] - ref["scores"][0, 1:]).abs()
    stable = torch.ones(mt["K"], dtype=torch.bool)
    stable[:-1] &= gaps > 1e-4
    stable[1:] &= gaps > 1e-4
    stable[-1] = False
    assert torch.equal(dec["xs"].cpu()[0][stable], ref["xs"][0][stable])
    assert torch.equal(dec["ys"].cpu()[0][stable], ref["ys"][0][stable])
    assert torch.equal(dec["clses"].cpu()[0][stable], ref["clses"][0][stable])


@pytest.mark.parametrize("tensor_core", [False, True], ids=["simt", "tc3x"])
def test_c1_512_frame_vs_oracle(dev, tensor_core):
    """BASELINE config C1: 512x512 frame, DLA-34+DCNv2 forward + 10x10 affinity, CPU oracle vs B200."""
    heads = spec.MOT_HEADS
    m, sd, opt = build(heads, "mot", dev, tensor_core=tensor_core)
    tol = TOL if tensor_core else TOL_ANCHOR
    x = spec.make_input((1, 3, 512, 512), tag="c1")
    torch.set_num_threads(oracle_threads())
    with torch.no_grad():
        ref_out, ref_fm = restate.dlaseg_forward(sd, x, heads)
    out, fm = m(x.to(dev))
    errs = {h: nerr(out[0][h].cpu(), ref_out[0][h]) for h in heads}
    for i in range(13):
        errs[f"fm{i}"] = nerr(fm[i].cpu(), ref_fm[i])
    centres = spec.make_uniform((1, 10, 1, 1, 2), -1, 1, tag="c1c")
    centres_prev = spec.make_uniform((1, 10, 1, 1, 2), -1, 1, tag="c1p")
    with torch.no_grad():
        e_ref = restate.afe_extract(sd, ref_fm, centres)
        p_ref = restate.afe_extract(sd, ref_fm, centres_prev)
        a_ref = restate.afe_affinity(sd, p_ref, e_ref, 100, False)
    e = m.AFE.forward_feature_extracter(fm, centres.to(dev))
    p = m.AFE.forward_feature_extracter(fm, centres_prev.to(dev))
    a = m.AFE.forward_stacker_features(p, e, False)
    errs["emb"] = nerr(e.cpu(), e_ref)
    errs["aff"] = nerr(a, a_ref)
    assert a.shape == (10, 11)
    print("C1", "tc" if tensor_core else "simt", {k: f"{v:.2e}" for k, v in errs.items()})
    bad = {k: v for k, v in errs.items() if not v < tol}
    assert not bad, bad
    # decode on the ORACLE's heat-map, shared tensor: bit-exact
    from deft_b200.model.decode import generic_decode
    sig = restate.sigmoid_output(ref_out[0])
    dref = restate.generic_decode(sig, K=100)
    dgpu = generic_decode({k: v.to(dev) for k, v in sig.items()}, K=100, opt=opt)
    for k in ("scores", "clses", "xs", "ys", "bboxes", "tracking"):
        assert torch.equal(dgpu[k].cpu(), dref[k]), k


def test_c2_full_size_properties(dev):
    """1088x608 (BASELINE C2), where the CPU oracle is too slow for the test suite: size-independent properties.
    (1) CUDA graph replay == eager launch list, bit for bit; (2) determinism; (3) FeatureMaps idx9 is idx5,
    idx10 equals idx6 (dla.py:792-799); (4) decode output sorted, idempotent NMS; (5) affinity rows are
    probabilities; (6) DCN linearity in the weights on a full-size layer."""
    heads = spec.MOT_HEADS
    m, sd, opt = build(heads, "mot", dev)
    x = spec.make_input((1, 3, 608, 1088), tag="c2").to(dev)
    out, fm = m(x)
    h1 = {k: v.clone() for k, v in out[0].items()}
    f1 = [f.clone() for f in fm]
    shapes = [(16, 608, 1088), (32, 304, 544), (64, 152, 272), (128, 76, 136), (256, 38, 68), (512, 19, 34), (64, 152, 272),
              (128, 76, 136), (256, 38, 68), (512, 19, 34), (64, 152, 272), (64, 152, 272), (64, 152, 272)]
    assert [tuple(f.shape[1:]) for f in fm] == shapes
    assert torch.equal(f1[9], f1[5]) and torch.equal(f1[10], f1[6])
    for f in f1:
        assert torch.isfinite(f).all()
    m.use_cuda_graph = True
    m.invalidate()
    out2, fm2 = m(x)
    out2, fm2 = m(x)
    for k in h1:
        assert torch.equal(out2[0][k], h1[k]), k
    for a, b in zip(fm2, f1):
        assert torch.equal(a, b)
    from deft_b200.model.decode import generic_decode, sigmoid_output
    o = sigmoid_output({k: v.clone() for k, v in h1.items()}, opt)
    dec = generic_decode(o, K=100, opt=opt)
    s = dec["scores"][0]
    assert (s[:-1] >= s[1:]).all() and (s > 0).all()
    flat = dec["ys"][0].long() * 272 + dec["xs"][0].long()
    assert flat.unique().numel() == 100
    assert torch.equal(o["hm"].view(-1)[flat], s)                     # scores are the heat values at the peaks
    centres = torch.stack([dec["xs"][0] / 272 * 2 - 1, dec["ys"][0] / 152 * 2 - 1], 1).view(1, 100, 1, 1, 2)
    emb = m.AFE.forward_feature_extracter(fm, centres)
    assert emb.shape == (1, 100, 416) and torch.isfinite(emb).all() and (emb >= 0).all()
    mem = torch.stack([emb[0] * (1 + 0.01 * f) for f in range(5)])   # 5 memory frames x 100 = 500 track embeddings
    aff = m.AFE.affinity_batched(mem.contiguous(), [100] * 5, emb[0].contiguous())
    assert aff.shape == (5, 100, 101)
    assert (aff >= 0).all() and (aff <= 1).all()
    # same-frame self affinity must differ from a perturbed memory frame (the MLP sees the memory)
    assert not torch.equal(aff[0], aff[4])
    # DCN linearity in W on the biggest DCN layer shape (64->64 @152x272)
    from deft_b200 import ops, packing
    sp = {}
    spec._deform(sp, "d", 64, 64)
    s1 = {k: v.to(dev) for k, v in spec.make_state_dict(sp, seed=1).items()}
    s2 = dict(s1)
    s2["d.conv.weight"] = spec.make_state_dict(sp, seed=2)["d.conv.weight"].to(dev)
    s3 = dict(s1)
    s3["d.conv.weight"] = s1["d.conv.weight"] + s2["d.conv.weight"]
    xin = ops.as_nhwc(f1[12])
    ys = [ops.dcnv2_forward(xin, packing.pack_dcn(s, "d.conv"), relu=False).nchw_view() for s in (s1, s2, s3)]
    bias = s1["d.conv.bias"].view(1, -1, 1, 1)
    assert nerr((ys[0] + ys[1] - bias).cpu(), ys[2].cpu()) < 1e-5


@pytest.mark.parametrize("graph", [False, True], ids=["eager", "graph"])
def test_frame_pipeline_depths_and_public_api_agree(dev, graph):
    """FramePipeline (the one-call-per-frame API bench.py times): (1) its detections for a frame are bit-identical
    to DLASeg.forward -> sigmoid_output -> generic_decode through the reference-named API; (2) its embeddings /
    affinity equal AFE.forward_feature_extracter / affinity_batched on the same memory; (3) the software-pipelined modes (depth 2: dense part of frame t+1 over the tail of frame t; depth 3: two dense
    parts in flight as well) return exactly what the sequential mode returns, frame by frame, over a sequence long enough to wrap the memory ring."""
    from deft_b200.model.decode import generic_decode, sigmoid_output
    from deft_b200.pipeline import FramePipeline
    heads = spec.MOT_HEADS
    H, W, K, F = 96, 160, 20, 3
    frames = [spec.make_input((1, 3, H, W), tag=f"pipe{i}").to(dev) for i in range(7)]
    results = {}
    for depth in (1, 2, 3, 4):                               # 4 = what bench.py runs
        m, sd, opt = build(heads, "mot", dev, max_object=32, K=K)
        pipe = FramePipeline(m, opt, H, W, mem_frames=F, device=dev, use_cuda_graph=graph, depth=depth)
        pipe.xs[0].copy_(frames[0])
        pipe.warmup()
        if graph:                                            # both halves of every arena are CUDA graphs
            assert all(p.graph is not None for p in pipe.plans) and all(g is not None for g in pipe.tail_graph)
        pipe.reset_memory()
        seq = []
        for x in frames:
            r = pipe.step_device(x)
            if depth > 1:                                    # deliberately no host sync between frames
                # results of frame t must be read before frame t+depth reuses the arena: snapshot on the tail stream
                with torch.cuda.stream(pipe.tail_stream):
                    snap = {k: r[k].clone() for k in ("scores", "clses", "bboxes", "embeddings", "affinity")}
            else:
                snap = {k: r[k].clone() for k in ("scores", "clses", "bboxes", "embeddings", "affinity")}
            seq.append(snap)
        pipe.synchronize()
        torch.cuda.synchronize(dev)
        results[depth] = (seq, m, opt)
    seq1, m, opt = results[1]
    for depth in (2, 3, 4):
        for t, (a, b) in enumerate(zip(seq1, results[depth][0])):
            for k in a:
                assert torch.equal(a[k], b[k]), (depth, t, k)
    # frame 0 against the piecewise public API (fresh memory = zeros)
    out, fm = m(frames[0])
    o = sigmoid_output({k: v.clone() for k, v in out[0].items()}, opt)
    dec = generic_decode(o, K=K, opt=opt)
    assert torch.equal(dec["scores"], seq1[0]["scores"]) and torch.equal(dec["bboxes"], seq1[0]["bboxes"])
    assert torch.equal(dec["clses"], seq1[0]["clses"])
    oh, ow = H // 4, W // 4
    bb = dec["bboxes"][0]
    # convert_detection (image.py:391-412) in the kernel's operation order
    cx = (2.0 * (bb[:, 0] / ow) + (bb[:, 2] - bb[:, 0]) / ow) - 1.0
    cy = (2.0 * (bb[:, 1] / oh) + (bb[:, 3] - bb[:, 1]) / oh) - 1.0
    centres = torch.stack([cx, cy], 1).view(1, K, 1, 1, 2)
    emb = m.AFE.forward_feature_extracter(fm, centres)
    assert nerr(emb[0].cpu(), seq1[0]["embeddings"].cpu()) < 1e-4
    # frame 0 met an empty recorder: every block is zero.  Frame 1: the recorder holds frame 0's K embeddings in ring
    # slot 0 (frame index 0), the other slots are still empty (no live rows -> zero blocks, not "zero embeddings")
    assert float(seq1[0]["affinity"].abs().max()) == 0.0
    mem = torch.zeros((F, 32, emb.shape[2]), device=dev)
    mem[0, :K] = seq1[0]["embeddings"]
    aff = m.AFE.affinity_batched(mem, [K] + [0] * (F - 1), seq1[1]["embeddings"].contiguous())
    assert aff.shape == seq1[1]["affinity"].shape
    assert nerr(aff[0, :K, : K + 1].cpu(), seq1[1]["affinity"][0, :K, : K + 1].cpu()) < 1e-4
    assert float(seq1[1]["affinity"][1:].abs().max()) == 0.0 and float(seq1[1]["affinity"][0, K:].abs().max()) == 0.0
    # frame F+1 wraps the ring: every slot is live
    assert float(seq1[F + 1]["affinity"][:, :K, :K].min()) > 0.0


@pytest.mark.parametrize("name,heads,dataset,hw,K,F,lstm_tracks", [
    ("kitti", spec.KITTI_HEADS, "kitti", (384, 1248), 50, 4, 200),        # BASELINE config C3: 1242x375 padded to /32, 50 dets x 200 tracks, LSTM step
    ("nuscenes", spec.NUSC_HEADS, "nuscenes", (928, 1600), 100, 5, 0),    # BASELINE config C4 per camera: 1600x900 padded to /32
])
def test_c3_c4_shapes_pipeline_properties(dev, name, heads, dataset, hw, K, F, lstm_tracks):
    """BASELINE configs C3 / C4 at full resolution (the CPU oracle is too slow there): the whole FramePipeline runs,
    every output is finite, decode is sorted with unique peaks, affinity entries are probabilities, and the
    pipelined mode (3 frames in flight) reproduces the sequential mode bit for bit on a 4-frame sequence."""
    from deft_b200.kalman_filter_lstm import DecoderRNN, KalmanFilterLSTM
    from deft_b200.pipeline import FramePipeline
    H, W = hw
    frames = [spec.make_input((1, 3, H, W), tag=f"{name}{i}").to(dev) for i in range(4)]
    seqs = {}
    for depth in (1, 3):
        m, sd, opt = build(heads, dataset, dev, max_object=100, K=K)
        kf = None
        if lstm_tracks:
            rnn = DecoderRNN(128, dataset)
            rnn.load_state_dict(spec.make_state_dict(spec.lstm_spec(dataset), seed=55))
            kf = KalmanFilterLSTM(opt, rnn.to(dev))
        pipe = FramePipeline(m, opt, H, W, mem_frames=F, lstm=kf, lstm_tracks=lstm_tracks, device=dev, depth=depth)
        if lstm_tracks:
            pipe.lstm_x.copy_(spec.make_input(tuple(pipe.lstm_x.shape), tag="lx").to(dev))
        pipe.xs[0].copy_(frames[0])
        pipe.warmup()
        pipe.reset_memory()
        if lstm_tracks:
            pipe.lstm_h.zero_()
            pipe.lstm_c.zero_()
        seq = []
        keys = ["scores", "clses", "bboxes", "embeddings", "affinity"] + (["lstm_pred"] if lstm_tracks else [])
        for x in frames:
            r = pipe.step_device(x)
            with torch.cuda.stream(pipe.tail_stream if pipe.tail_stream is not None else torch.cuda.current_stream(dev)):
                seq.append({k: r[k].clone() for k in keys})
        pipe.synchronize()
        torch.cuda.synchronize(dev)
        seqs[depth] = seq
        del pipe, m
        torch.cuda.empty_cache()
    for t, (a, b) in enumerate(zip(seqs[1], seqs[3])):
        for k in a:
            assert torch.equal(a[k], b[k]), (t, k)
    last = seqs[1][-1]
    for k, v in last.items():
        assert torch.isfinite(v.float()).all(), k
    s = last["scores"][0]
    assert s.shape[0] == K and (s[:-1] >= s[1:]).all() and (s > 0).all()
    aff = last["affinity"][:, :K, : K + 1]
    assert (aff >= 0).all() and (aff <= 1).all()
    assert last["embeddings"].shape == (K, 704 if dataset == "nuscenes" else 416)
