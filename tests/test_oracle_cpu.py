"""CPU suite, part 1: the oracle restatement is pinned to tests/golden (outputs of the REAL reference
modules, oracle/gen_golden.py) and, when /root/reference is present, re-checked live against it."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, nerr
from oracle import ref_loader, restate, spec

TOL = 2e-5   # two fp32 CPU evaluations of the same formulas (different op order)

FRAMES = [("frame_mot_96x128", spec.MOT_HEADS, "mot"), ("frame_kitti_64x160", spec.KITTI_HEADS, "kitti"),
          ("frame_nusc_64x96", spec.NUSC_HEADS, "nuscenes")]


@pytest.mark.parametrize("name,heads,dataset", FRAMES)
def test_state_dict_inventory(name, heads, dataset):
    with open(os.path.join(GOLDEN, f"state_dict_keys_{dataset}_{len(heads)}heads.json")) as f:
        ref = {k: tuple(v) for k, v in json.load(f).items()}
    mine = {k: tuple(s) for k, (_, s) in spec.dlaseg_spec(heads, dataset=dataset).items()}
    assert mine == ref


@pytest.mark.parametrize("name,heads,dataset", FRAMES)
@pytest.mark.parametrize("explicit", [False, True])
def test_frame_restatement_vs_reference_golden(name, heads, dataset, explicit):
    g = torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)
    mt = g["meta"]
    sd = spec.make_state_dict(spec.dlaseg_spec(heads, dataset=dataset))
    x = spec.make_input((1, 3, mt["H"], mt["W"]), tag="image")
    with torch.no_grad():
        out, fm = restate.dlaseg_forward(sd, x, heads, explicit=explicit)
    for h in heads:
        assert nerr(out[0][h], g["heads"][h]) < TOL, h
    assert len(fm) == 13
    for i, (f, d) in enumerate(zip(fm, g["fm"])):
        assert list(f.shape) == d["shape"]
        assert nerr(f.contiguous().view(-1)[d["pos"]], d["val"]) < TOL * 5, i
        assert abs(float(f.double().abs().sum()) - d["abssum"]) / d["abssum"] < 1e-5
    centres = spec.make_uniform((1, mt["N"], 1, 1, 2), -1.05, 1.05, tag="centres")
    centres_prev = spec.make_uniform((1, mt["n_prev"], 1, 1, 2), -1.0, 1.0, tag="centres_prev")
    with torch.no_grad():
        emb = restate.afe_extract(sd, fm, centres)
        emb_prev = restate.afe_extract(sd, fm, centres_prev)
    assert nerr(emb, g["emb"]) < TOL * 5
    assert nerr(emb_prev, g["emb_prev"]) < TOL * 5
    P = mt["max_object"]
    with torch.no_grad():
        assert nerr(restate.afe_affinity(sd, g["emb_prev"], g["emb"], P, False), g["aff"]) < 1e-5
        assert nerr(restate.afe_affinity(sd, g["emb_prev"], g["emb"], P, True), g["aff_fill"]) < 1e-5
        assert nerr(restate.afe_affinity(sd, g["emb"], g["emb"], P, False), g["aff_full"]) < 1e-5
    # decode on the reference's own (sigmoided) heads: bit-exact
    sig = restate.sigmoid_output(g["heads"])
    dec = restate.generic_decode(sig, K=mt["K"])
    for k, v in g["decode"].items():
        assert torch.equal(dec[k], v), k


def test_decode_plateau_ties_bit_exact():
    g = torch.load(os.path.join(GOLDEN, "decode_plateau.pt"), weights_only=False)
    dec = restate.generic_decode(g["in"], K=g["K"])
    # scores and boxes must be bit-exact; among exactly tied scores torch.topk's order is
    # implementation-defined, so indices are compared as per-score multisets
    assert torch.equal(dec["scores"], g["decode"]["scores"])
    for b in range(dec["scores"].shape[0]):
        mine = sorted(zip(dec["scores"][b].tolist(), dec["clses"][b].tolist(), dec["ys"][b].tolist(), dec["xs"][b].tolist()))
        ref = sorted(zip(g["decode"]["scores"][b].tolist(), g["decode"]["clses"][b].tolist(),
                         g["decode"]["ys"][b].tolist(), g["decode"]["xs"][b].tolist()))
        # strictly-greater-than-threshold part must agree exactly
        thr = dec["scores"][b].min().item()
        assert [m for m in mine if m[0] > thr] == [r for r in ref if r[0] > thr]


def test_dcn_layers_vs_reference_golden():
    cases = torch.load(os.path.join(GOLDEN, "dcn_layers.pt"), weights_only=False)
    for c in cases:
        sp = {}
        spec._deform(sp, "d", c["chi"], c["cho"])
        sd = spec.make_state_dict(sp, seed=c["seed"])
        x = spec.make_input((1, c["chi"], c["h"], c["w"]), tag=f"dcn{c['chi']}x{c['cho']}")
        with torch.no_grad():
            for explicit in (False, True):
                y = restate.deform_conv(sd, "d", x, explicit=explicit)
                assert nerr(y, c["y"]) < TOL, (c["chi"], c["cho"], explicit)


def test_dcn_identities():
    """Known-answer identities of SURVEY.md section 4: zero offset/mask conv => 0.5*conv2d+b; integer
    offsets + mask=1 => shifted conv."""
    g = torch.Generator(); g.manual_seed(3)
    x = torch.randn(1, 8, 9, 11, generator=g)
    w = torch.randn(6, 8, 3, 3, generator=g) * 0.2
    b = torch.randn(6, generator=g)
    off = torch.zeros(1, 18, 9, 11)
    m = torch.full((1, 9, 9, 11), 0.5)
    y = restate.dcn_v2_explicit(x, off, m, w, b)
    ref = 0.5 * torch.nn.functional.conv2d(x, w, None, padding=1) + b.view(1, -1, 1, 1)
    assert nerr(y, ref) < 1e-6
    off[:, 0::2] = 1.0      # dy = +1 on every tap -> one-row shift
    y = restate.dcn_v2_explicit(x, off, torch.ones(1, 9, 9, 11), w, b)
    # rows shifted by one: sample(h-1+ky+1) ; zero beyond the border (row H is outside)
    xpad = torch.nn.functional.pad(x, (1, 1, 1, 2))     # rows -1..H+1
    ref = torch.nn.functional.conv2d(xpad[:, :, 1:], w, None) + b.view(1, -1, 1, 1)
    assert nerr(y, ref) < 1e-6


def test_lstm_restatement_vs_torch_lstm_golden():
    g = torch.load(os.path.join(GOLDEN, "lstm_step.pt"), weights_only=False)
    for dataset, fin, fut in (("mot", 11, 5), ("nuscenes", 18, 4)):
        sd = spec.make_state_dict(spec.lstm_spec(dataset), seed=55)
        T = g[dataset]["T"]
        h0 = spec.make_input((T, 128), tag="h0") * 0.1
        c0 = spec.make_input((T, 128), tag="c0") * 0.1
        x = spec.make_input((T, fin), tag="lstm_x")
        hn, cn, pred = restate.lstm_predict(sd, h0, c0, x, fut)
        assert nerr(hn, g[dataset]["hn"]) < 1e-6
        assert nerr(cn, g[dataset]["cn"]) < 1e-6
        assert nerr(pred, g[dataset]["pred"]) < 1e-5


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree only exists in the build container")
def test_live_reference_affinity_depends_on_max_object():
    """The dual softmax denominators include (max_object - n) zero entries (AFE.py:123-134)."""
    sd = spec.make_state_dict(spec.dlaseg_spec(spec.MOT_HEADS))
    xp = spec.make_input((1, 5, 416), tag="xp").abs()
    xn = spec.make_input((1, 6, 416), tag="xn").abs()
    with torch.no_grad():
        a = restate.afe_affinity(sd, xp, xn, 100, False)
        b = restate.afe_affinity(sd, xp, xn, 200, False)
        _, AFE, _ = ref_loader.load()
        mod = AFE.AFE_module(ref_loader.make_opt(max_object=200)).eval()
        mod.load_state_dict({k[4:]: v for k, v in sd.items() if k.startswith("AFE.")})
        ref = mod.forward_stacker_features(xp, xn, False)
    assert np.abs(a - b).max() > 1e-3
    assert nerr(b, ref) < 1e-5


def test_afe_train_forward_and_loss_restatement_vs_reference_golden():
    """Row a9: restate.afe_forward / afe_loss against the REAL AFE_module.forward / .loss (golden produced by
    oracle/gen_golden.py::gen_afe_train in the build container)."""
    from oracle.gen_golden import afe_train_inputs
    g = torch.load(os.path.join(GOLDEN, "afe_train_forward.pt"), weights_only=False)
    B, P = g["meta"]["B"], g["meta"]["P"]
    sd = spec.make_state_dict(spec.dlaseg_spec(spec.MOT_HEADS))
    fm_pre, fm_next, l_pre, l_next, target, mask0, mask1 = afe_train_inputs(B, P)
    with torch.no_grad():
        x = restate.afe_forward(sd, fm_pre, fm_next, l_pre, l_next, P)
        vals, idx = restate.afe_loss(g["x"], target, mask0, mask1, P)
    assert nerr(x, g["x"]) < 2e-5
    assert np.allclose(vals, g["loss"], rtol=1e-6, atol=1e-7)
    assert torch.equal(idx, g["indexes_pre"])


def test_track_similarity_restatement_properties():
    """Row f1 oracle (the tracker module itself cannot be imported: SURVEY 8(c)): the restated get_similarity keeps the
    reference's row-selection quirk (exactly mm+1 rows -> all of them, more -> the last mm) and its decay rule."""
    rng = np.random.RandomState(0)
    sims = {f: rng.rand(5, 8).astype(np.float32) for f in range(10)}
    rec = restate.recorder_similarity(sims, 10, dataset="mot")
    assert all(np.array_equal(rec[f], sims[f]) for f in range(1, 10))                        # decay = 1.0 inside m_frame = 10
    assert np.array_equal(rec[0], (sims[0] * pow(0.01, 10 / 3.0)).astype(np.float32))         # beyond: decay2 ** (d/3)
    nodes = [(f, f % 5) for f in range(3, 10)]
    rows = np.array([rec[f][i] for f, i in nodes])
    assert np.array_equal(restate.track_similarity(nodes[-4:], 10, rec), np.median(rows[-4:], axis=0))
    assert np.array_equal(restate.track_similarity(nodes[-5:], 10, rec), np.median(rows[-5:], axis=0))   # mm+1 rows: all five
    assert np.array_equal(restate.track_similarity(nodes, 10, rec), np.median(rows[-4:], axis=0))        # seven rows: the last four
    assert restate.track_similarity([], 10, rec) is None
    assert np.array_equal(restate.track_similarity(nodes[-3:], 10, rec, dataset="nuscenes"), np.median(rows[-3:], axis=0))
    assert np.array_equal(restate.track_similarity(nodes[-4:], 10, rec, dataset="nuscenes"), np.median(rows[-2:], axis=0))


def test_tracker_restatements_vs_the_reference_tracker_code_golden():
    """Rows f1 / f3 pinned: tests/golden/tracker_logic.pt holds outputs of the reference's OWN FeatureRecorder.update,
    STrack.get_similarity (utils/tracker.py:46-90, 219-252) and matching.fuse_motion + KalmanFilterLSTM.gating_distance
    (matching.py:311-364, kalman_filter_lstm.py:80-103) -- the definitions are cut out of the reference source with `ast` and run
    unmodified by oracle/gen_golden.py::gen_tracker_logic (the modules themselves do not import).
    One environment effect: `pre_similarity * delta` is float32-array x numpy-float64-scalar (delta inherits np.int64 frame
    indices).  The reference's numpy 1.x (environment.yml) keeps that product float32 (value-based casting); the numpy 2 of this
    container promotes it to float64 (NEP 50), so the golden holds f64(sim) * delta.  The restatement (and the device kernel)
    follow the reference's environment: float32.  Hence: decayed similarities within one float32 rounding of the golden and
    bit-equal where delta = 1; the per-track row selection + medians bit-exact on the golden's own similarities; fuse_motion
    (float64 on both sides) bit-exact."""
    g = torch.load(os.path.join(GOLDEN, "tracker_logic.pt"), weights_only=False)
    assert g["constants"] == {"Max_record_frame": 50, "decay": 1.0, "decay2": 0.01, "max_track_node": 50}
    n_tracks = 0
    for dataset, rec in g["recorder"].items():
        frames = rec["frames"]
        for chk in rec["checks"]:
            fi = chk["frame"]
            on_record = frames[max(0, frames.index(fi) - 49): frames.index(fi)]              # <= 50 frames on record incl. the current one
            assert sorted(chk["all_similarity"]) == on_record
            sim = restate.recorder_similarity(chk["aff"], fi, dataset=dataset)
            assert sorted(sim) == on_record
            m_frame = {"mot": 10, "kitti_tracking": 5, "nuscenes": 3}[dataset]
            for pre in on_record:
                ref = chk["all_similarity"][pre]
                assert sim[pre].dtype == np.float32
                if fi - pre < m_frame:                                                       # decay = 1.0: the product is the affinity itself
                    assert np.array_equal(sim[pre], ref), (dataset, fi, pre)
                else:
                    assert np.array_equal(sim[pre], (chk["aff"][pre] * np.float32(pow(0.01, (fi - pre) / 3.0))).astype(np.float32))
                    assert np.all(np.abs(sim[pre].astype(np.float64) - ref) <= 1.3e-7 * np.abs(ref) + 2e-38), (dataset, fi, pre)   # delta and the product each round once to float32; below 1.2e-38 float32 is denormal
            for tr in chk["tracks"]:
                got = restate.track_similarity(tr["nodes"], fi, chk["all_similarity"], dataset=dataset)
                assert (got is None) == (tr["similarity"] is None)
                if got is not None:
                    assert np.array_equal(got, tr["similarity"]), (dataset, fi, tr["nodes"])
                    f32 = restate.track_similarity(tr["nodes"], fi, sim, dataset=dataset)     # the float32 chain the device kernel mirrors
                    assert np.allclose(f32, tr["similarity"], rtol=1e-6, atol=1e-30)
                    n_tracks += 1
    assert n_tracks > 150
    for case in g["fuse_motion"]:
        got = restate.fuse_motion_lstm(case["cost"], case["pred"][:, :2], case["n_obs"], case["det"][:, :2])
        assert np.array_equal(np.isinf(got), np.isinf(case["fused"]))
        assert np.array_equal(got[np.isfinite(got)], case["fused"][np.isfinite(case["fused"])])
        assert np.isinf(case["fused"]).any() or case["cost"].shape[0] == 1


def test_preprocess_restatement_vs_cv2():
    """Row f2 oracle: the restated bilinear warp stays within cv2.warpAffine's 1/32-pixel fixed-point error (the
    reference calls cv2: detector.py:399-407); the normalisation is detector.py:408-410 verbatim."""
    cv2 = pytest.importorskip("cv2")
    g = torch.Generator().manual_seed(5)
    src = torch.randint(0, 256, (70, 110, 3), generator=g, dtype=torch.uint8).numpy()
    # smooth the source: the fixed-point coordinate error is bounded per unit of gradient
    src = cv2.GaussianBlur(src, (7, 7), 2.0)
    fwd = np.array([[1.3, 0.05, 4.0], [-0.03, 1.25, -2.0]], dtype=np.float32)               # trans_input: source -> network input
    inv = cv2.invertAffineTransform(fwd)
    ref = cv2.warpAffine(src, fwd, (96, 64), flags=cv2.INTER_LINEAR).astype(np.float64)
    got = restate.warp_affine_bilinear(src, inv, 64, 96)
    inner = (got > 0).all(axis=2)
    assert np.abs(got - ref)[inner].max() <= 4.0                                             # of 255: coordinate quantisation x gradient
    assert np.abs(got - ref)[inner].mean() < 0.6
    x = restate.normalize_u8(src, [0.408, 0.447, 0.470], [0.289, 0.274, 0.278])
    assert x.shape == (1, 3, 70, 110) and x.dtype == torch.float32
    assert abs(float(x[0, 1, 3, 4]) - ((src[3, 4, 1] / 255.0 - 0.447) / 0.274)) < 1e-6


def test_linear_assignment_restatement_is_optimal_on_small_cases():
    """Row f3 oracle: the restated lapjv(extend_cost, cost_limit) equals the brute-force optimum of
    sum(matched c) + limit/2 * (#unmatched rows + #unmatched cols) on every small matrix tried."""
    import itertools
    rng = np.random.RandomState(3)
    for n, m in ((3, 3), (4, 2), (2, 5), (5, 4)):
        for _ in range(6):
            cost = rng.rand(n, m)
            thresh = 0.6
            x, y = restate.linear_assignment(cost, thresh)
            best, arg = None, None
            cols = list(range(m)) + [-1] * n
            for perm in set(itertools.permutations(cols, n)):
                used = [j for j in perm if j >= 0]
                if len(set(used)) != len(used):
                    continue
                k = len(used)
                tot = sum(cost[i, j] for i, j in enumerate(perm) if j >= 0) + thresh / 2 * ((n - k) + (m - k))
                if best is None or tot < best - 1e-12:
                    best, arg = tot, perm
            assert list(x) == list(arg), (cost, x, arg)
            assert all(cost[i, j] <= thresh for i, j in enumerate(x) if j >= 0)
    fused = restate.fuse_motion_lstm(np.array([[0.5, 0.2], [0.3, 0.4]]), np.array([[0., 0.], [10., 10.]]), [300, 5], np.array([[1., 1.], [9., 9.]]))
    assert np.isinf(fused[0, 1]) and abs(fused[0, 0] - (0.45 + 0.005 * 2.0)) < 1e-12 and np.allclose(fused[1], [0.27, 0.36])
