"""Generate tests/golden/* by executing the REAL reference modules (build container only).

    python -m oracle.gen_golden

TEST INFRASTRUCTURE ONLY.  The reference holds no golden vectors of its own, so
these fixtures are "outputs of the reference itself run here": the reference's DLASeg / AFE_module /
generic_decode (imported from /root/reference/src/lib through oracle/ref_loader.py, DCN =
torchvision.ops.deform_conv2d shim) and torch.nn.LSTM/Linear assembled as
utils/tracking_utils/kalman_filter_lstm.py:9-29 does (that file itself is un-importable: it
imports the broken model.model).  Weights/inputs are the seeded synthetic tensors of oracle/spec.py, so
the GPU box regenerates identical bits without the reference.  The tracker-side logic (rows f1 / f3: FeatureRecorder,
STrack.get_similarity, matching.fuse_motion) lives in un-importable modules too; gen_tracker_logic() cuts those definitions out
of the reference source with `ast` and executes them unmodified (tests/golden/tracker_logic.pt).
"""
import json
import os
import sys

import numpy as np
import torch

from . import ref_loader, spec

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def sample_positions(numel, n=512, tag="pos"):
    g = torch.Generator()
    g.manual_seed(12345 + len(tag) + numel)
    return torch.randint(0, numel, (n,), generator=g)


def fm_digest(t):
    flat = t.contiguous().view(-1)
    pos = sample_positions(flat.numel())
    return {"shape": list(t.shape), "sum": float(flat.double().sum()), "abssum": float(flat.double().abs().sum()),
            "max": float(flat.abs().max()), "pos": pos, "val": flat[pos].clone()}


def gen_frame(name, heads, dataset, H, W, N, n_prev, max_object, K):
    m, opt = ref_loader.build_dlaseg(heads, dataset=dataset, max_object=max_object)
    sp = spec.dlaseg_spec(heads, dataset=dataset)
    sd = spec.make_state_dict(sp)
    m.load_state_dict(sd, strict=True)
    x = spec.make_input((1, 3, H, W), tag="image")
    centres = spec.make_uniform((1, N, 1, 1, 2), -1.05, 1.05, tag="centres")      # some beyond the border
    centres_prev = spec.make_uniform((1, n_prev, 1, 1, 2), -1.0, 1.0, tag="centres_prev")
    _, _, decode = ref_loader.load()
    with torch.no_grad():
        out, fm = m(x)
        emb = m.AFE.forward_feature_extracter(fm, centres)
        emb_prev = m.AFE.forward_feature_extracter(fm, centres_prev)
        aff = m.AFE.forward_stacker_features(emb_prev, emb, False)
        aff_fill = m.AFE.forward_stacker_features(emb_prev, emb, True)
        aff_full = m.AFE.forward_stacker_features(emb, emb, False)
        sig = {k: v.clone() for k, v in out[0].items()}
        sig["hm"] = sig["hm"].sigmoid_()
        if "dep" in sig:
            sig["dep"] = 1.0 / (sig["dep"].sigmoid() + 1e-6) - 1.0
        dec = decode.generic_decode({k: v.clone() for k, v in sig.items()}, K=K, opt=opt)
    g = {"meta": dict(heads=dict(heads), dataset=dataset, H=H, W=W, N=N, n_prev=n_prev, max_object=max_object, K=K, seed=317),
         "heads": {k: v.clone() for k, v in out[0].items()},
         "fm": [fm_digest(f) for f in fm],
         "emb": emb.clone(), "emb_prev": emb_prev.clone(),
         "aff": torch.from_numpy(aff.copy()), "aff_fill": torch.from_numpy(aff_fill.copy()),
         "aff_full": torch.from_numpy(aff_full.copy()),
         "decode": {k: v.clone() for k, v in dec.items()}}
    torch.save(g, os.path.join(OUT, name + ".pt"))
    keys = {k: list(v.shape) for k, v in m.state_dict().items()}
    with open(os.path.join(OUT, f"state_dict_keys_{dataset}_{len(heads)}heads.json"), "w") as f:
        json.dump(keys, f, indent=0, sort_keys=True)
    print(name, "ok;", len(keys), "keys")


def gen_dcn():
    """One reference DeformConv module (dla.py:646-665) per (chi, cho, h, w) at small sizes."""
    dla, _, _ = ref_loader.load()
    cases = []
    for (chi, cho, h, w) in ((64, 64, 24, 40), (128, 64, 12, 20), (512, 256, 5, 9), (256, 128, 9, 7)):
        mod = dla.DeformConv(chi, cho).eval()
        sp = {}
        spec._deform(sp, "d", chi, cho)
        sd = spec.make_state_dict(sp, seed=1000 + chi + cho)
        mod.load_state_dict({k[2:]: v for k, v in sd.items()}, strict=True)
        x = spec.make_input((1, chi, h, w), tag=f"dcn{chi}x{cho}")
        with torch.no_grad():
            y = mod(x)
        cases.append({"chi": chi, "cho": cho, "h": h, "w": w, "seed": 1000 + chi + cho, "y": y.clone()})
    torch.save(cases, os.path.join(OUT, "dcn_layers.pt"))
    print("dcn ok")


def gen_decode_plateau():
    """Adversarial decode: plateaus/ties.  Records what the reference's generic_decode (torch.topk on
    this build) returns so the tie rule is documented by data."""
    _, _, decode = ref_loader.load()
    opt = ref_loader.make_opt()
    g = torch.Generator(); g.manual_seed(99)
    B, C, H, W, K = 2, 3, 16, 24, 12
    # heat quantised to 1/8 steps => many exact ties between distinct isolated peaks
    hm = (torch.rand(B, C, H, W, generator=g) * 8).floor() / 8 * 0.9 + 0.05
    out = {"hm": hm.clone(), "reg": torch.rand(B, 2, H, W, generator=g), "wh": torch.randn(B, 2, H, W, generator=g),
           "tracking": torch.randn(B, 2, H, W, generator=g), "ltrb_amodal": torch.randn(B, 4, H, W, generator=g)}
    with torch.no_grad():
        dec = decode.generic_decode({k: v.clone() for k, v in out.items()}, K=K, opt=opt)
    torch.save({"in": out, "K": K, "decode": dec}, os.path.join(OUT, "decode_plateau.pt"))
    print("decode plateau ok")


def gen_lstm():
    res = {}
    for dataset, fin, fout, fut in (("mot", 11, 20, 5), ("nuscenes", 18, 16, 4)):
        lstm = torch.nn.LSTM(fin, 128)                      # kalman_filter_lstm.py:13-20
        out1 = torch.nn.Linear(128, 64)
        out2 = torch.nn.Linear(64, fout)
        sd = spec.make_state_dict(spec.lstm_spec(dataset), seed=55)
        lstm.load_state_dict({k[5:]: v for k, v in sd.items() if k.startswith("lstm.")})
        out1.load_state_dict({"weight": sd["out1.weight"], "bias": sd["out1.bias"]})
        out2.load_state_dict({"weight": sd["out2.weight"], "bias": sd["out2.bias"]})
        T = 7
        h0 = spec.make_input((T, 128), tag="h0") * 0.1
        c0 = spec.make_input((T, 128), tag="c0") * 0.1
        x = spec.make_input((T, fin), tag="lstm_x")
        hn, cn, pred = [], [], []
        with torch.no_grad():
            for t in range(T):                              # one track at a time, like tracker.py:467
                nf = x[t].view(1, 1, fin).permute(1, 0, 2)  # predict(): permute(1,0,2)
                o, (h, c) = lstm(nf, (h0[t].view(1, 1, 128), c0[t].view(1, 1, 128)))
                y = out2(out1(o[-1]))
                hn.append(h.view(128)); cn.append(c.view(128)); pred.append(y.view(fut, -1))
        res[dataset] = {"T": T, "hn": torch.stack(hn), "cn": torch.stack(cn), "pred": torch.stack(pred)}
    torch.save(res, os.path.join(OUT, "lstm_step.pt"))
    print("lstm ok")


def afe_train_inputs(B, P, dataset="mot", seed_tag="a9"):
    """Seeded inputs of the training-side AFE surface: 13 FeatureMaps per frame pair (small maps, reference channel
    list AFE.py:15), centre tensors with 1.5-padded rows (image.py:356-359), a one-to-one label matrix with
    false-object assignments and the matching uint8 masks (generic_dataset.py builds them the same way)."""
    chans = spec.SELECTOR_IN
    sizes = [(24, 40), (12, 20), (6, 10), (3, 5), (2, 3), (1, 2), (6, 10), (3, 5), (2, 3), (1, 2), (6, 10), (6, 10), (6, 10)]
    fm_pre = [spec.make_input((B, c, h, w), tag=f"{seed_tag}_pre{k}").abs() for k, (c, (h, w)) in enumerate(zip(chans, sizes))]
    fm_next = [spec.make_input((B, c, h, w), tag=f"{seed_tag}_next{k}").abs() for k, (c, (h, w)) in enumerate(zip(chans, sizes))]
    l_pre = spec.make_uniform((B, P, 1, 1, 2), -1, 1, tag=f"{seed_tag}_lpre")
    l_next = spec.make_uniform((B, P, 1, 1, 2), -1, 1, tag=f"{seed_tag}_lnext")
    g = torch.Generator(); g.manual_seed(4242)
    target = torch.zeros(B, 1, P + 1, P + 1)
    mask0 = torch.zeros(B, 1, P + 1, dtype=torch.uint8)
    mask1 = torch.zeros(B, 1, P + 1, dtype=torch.uint8)
    for b in range(B):
        n_pre, n_next = P // 3 + 5 * b + 1, P // 2 - 3 * b
        l_pre[b, n_pre:] = 1.5
        l_next[b, n_next:] = 1.5
        mask0[b, 0, :n_pre] = 1; mask0[b, 0, P] = 1
        mask1[b, 0, :n_next] = 1; mask1[b, 0, P] = 1
        perm = torch.randperm(n_next, generator=g)
        for i in range(n_pre):                       # two thirds of the previous objects survive, the rest disappear
            if i < len(perm) and i % 3 != 2:
                target[b, 0, i, perm[i]] = 1
            else:
                target[b, 0, i, P] = 1
        matched = target[b, 0, :n_pre, :n_next].sum(0)
        for j in range(n_next):                      # unmatched next objects are new: row P
            if matched[j] == 0:
                target[b, 0, P, j] = 1
    return fm_pre, fm_next, l_pre, l_next, target, mask0, mask1


def gen_afe_train():
    """REAL AFE_module.forward (AFE.py:74-86) and .loss (235-328), eval mode, B=2, P=100."""
    _, AFE, _ = ref_loader.load()
    B, P = 2, 100
    opt = ref_loader.make_opt("mot", P)
    mod = AFE.AFE_module(opt).eval()
    sd_all = spec.make_state_dict(spec.dlaseg_spec(spec.MOT_HEADS))
    mod.load_state_dict({k[4:]: v for k, v in sd_all.items() if k.startswith("AFE.")}, strict=True)
    fm_pre, fm_next, l_pre, l_next, target, mask0, mask1 = afe_train_inputs(B, P)
    with torch.no_grad():
        x = mod(fm_pre, fm_next, l_pre, l_next)
        res = mod.loss(x, target, mask0.bool(), mask1.bool())     # bool masks: torch 2.x no longer indexes with uint8
    g = {"meta": dict(B=B, P=P), "x": x.clone(), "loss": [float(v) for v in res[:7]], "indexes_pre": res[7].clone()}
    torch.save(g, os.path.join(OUT, "afe_train_forward.pt"))
    print("afe train forward/loss ok:", g["loss"])


# --------------------------------------------------------------------------- f1 / f3: the tracker's own code, extracted
# utils/tracker.py and utils/matching.py cannot be imported here (they need `lap`, `cython_bbox`, `progress`, `nuscenes`, and
# tracker.py parses argv at import).  The pieces the product restates are plain numpy though, so they are cut out of the
# reference SOURCE with `ast` (class / function definitions and the module constants they read, nothing rewritten) and
# executed as they stand: FeatureRecorder + Node (tracker.py:29-137), STrack.get_similarity (219-252), matching.fuse_motion
# (matching.py:311-364), KalmanFilterLSTM.gating_distance (kalman_filter_lstm.py:80-103), chi2inv95 (kalman_filter.py:9-19).
REF_LIB = "/root/reference/src/lib"


def _extract(path, top_names=(), class_methods=()):
    """Source of the named top-level assignments / classes / functions of `path`, plus `Class.method` definitions re-rooted as
    top-level functions -- compiled from the unmodified AST nodes."""
    import ast
    tree = ast.parse(open(path).read())
    picked = []
    for node in tree.body:
        if isinstance(node, ast.Assign) and any(isinstance(t, ast.Name) and t.id in top_names for t in node.targets):
            picked.append(node)
        elif isinstance(node, (ast.ClassDef, ast.FunctionDef)) and node.name in top_names:
            picked.append(node)
        elif isinstance(node, ast.ClassDef):
            for sub in node.body:
                if isinstance(sub, ast.FunctionDef) and (node.name, sub.name) in class_methods:
                    picked.append(sub)
    mod = ast.Module(body=picked, type_ignores=[])
    return compile(ast.fix_missing_locations(mod), path, "exec")


def load_tracker_logic():
    """Namespace with the reference's own FeatureRecorder, Node, get_similarity, fuse_motion, gating_distance."""
    import types
    import scipy.linalg
    ns = {"np": np, "scipy": __import__("scipy")}
    exec(_extract(os.path.join(REF_LIB, "utils", "tracker.py"), ("Max_record_frame", "decay", "decay2", "max_track_node", "Node", "FeatureRecorder"),
                  (("STrack", "get_similarity"),)), ns)
    kf_ns = {}
    exec(_extract(os.path.join(REF_LIB, "utils", "tracking_utils", "kalman_filter.py"), ("chi2inv95",)), kf_ns)
    ns["kalman_filter"] = types.SimpleNamespace(chi2inv95=kf_ns["chi2inv95"])
    exec(_extract(os.path.join(REF_LIB, "utils", "matching.py"), ("fuse_motion",)), ns)
    exec(_extract(os.path.join(REF_LIB, "utils", "tracking_utils", "kalman_filter_lstm.py"), (), (("KalmanFilterLSTM", "gating_distance"),)), ns)
    return ns


def tracker_sequence(dataset, seed):
    """A synthetic video for the recorder: frame indices with gaps (so both decay branches and the 50-frame eviction run),
    2..7 detections per frame, and a stub model whose forward_stacker_features returns seeded float32 affinities."""
    rng = np.random.RandomState(seed)
    frames, f = [], 0
    for k in range(58):
        frames.append(f)
        f += 1 if rng.rand() < 0.8 else int(rng.randint(2, 5))
    counts = {fr: int(rng.randint(2, 6)) for fr in frames}
    return frames, counts


def gen_tracker_logic():
    import types
    ns = load_tracker_logic()
    out = {"constants": {k: ns[k] for k in ("Max_record_frame", "decay", "decay2", "max_track_node")}, "recorder": {}, "fuse_motion": []}
    for dataset, seed in (("mot", 11), ("kitti_tracking", 12), ("nuscenes", 13)):
        frames, counts = tracker_sequence(dataset, seed)
        rng = np.random.RandomState(seed + 100)
        aff = {}                                     # (frame, pre_frame) -> what the stub model returned

        class StubAFE:
            def forward_stacker_features(self, prev, cur, fill_up_column=False):
                assert fill_up_column is False
                a = rng.rand(len(prev), len(cur) + 1).astype(np.float32)
                aff[(int(cur[0, 0]), int(prev[0, 0]))] = a
                return a
        model = types.SimpleNamespace(AFE=StubAFE())
        rec = ns["FeatureRecorder"](dataset)
        checks = []
        trng = np.random.RandomState(seed + 200)
        for fi in frames:
            feats = np.full((counts[fi], 4), fi, dtype=np.float32)          # row 0 col 0 carries the frame index for the stub
            rec.update(model, fi, feats, np.zeros((counts[fi], 4), dtype=np.float32))
            pos = frames.index(fi)
            if not (pos in (3, 9) or pos % 8 == 7 or pos >= len(frames) - 2):    # a dozen check frames incl. before / after the eviction starts
                continue
            sim = {int(k): v.copy() for k, v in rec.all_similarity[fi].items()}
            # tracks: node lists (oldest first) over frames still on record, incl. single-node and > 5-node tracks and nodes
            # older than max_track_node frames are impossible here (record <= 50 frames) -- the skip branch is covered by
            # the restatement's own property test
            tracks = []
            on_record = [int(x) for x in rec.all_frame_index[:-1]]
            for _ in range(6):
                k = int(trng.randint(1, min(9, len(on_record)) + 1))
                fs = sorted(trng.choice(on_record, size=k, replace=False).tolist())
                nodes = [(f_, int(trng.randint(0, counts[f_]))) for f_ in fs]
                stub = types.SimpleNamespace(nodes=[ns["Node"](f_, i_) for f_, i_ in nodes], dataset=dataset)
                res = ns["get_similarity"](stub, fi, rec)
                tracks.append({"nodes": nodes, "similarity": None if res is None else np.asarray(res).copy()})
            checks.append({"frame": fi, "aff": {int(pre): aff[(fi, int(pre))].copy() for pre in sim}, "all_similarity": sim, "tracks": tracks})
        out["recorder"][dataset] = {"frames": frames, "counts": counts, "checks": checks}
    # fuse_motion (use_lstm=True as tracker.py:906-913 calls it; lambda_ default), both observation-count branches
    rng = np.random.RandomState(77)
    for T, m in ((6, 5), (1, 3), (9, 12)):
        cost = rng.rand(T, m).astype(np.float64)
        pred = rng.rand(T, 4) * np.array([40.0, 30.0, 1.0, 1.0])
        n_obs = [int(x) for x in rng.choice([1, 5, 299, 300, 301, 1000], size=T)]
        det = rng.rand(m, 4) * np.array([40.0, 30.0, 1.0, 50.0])
        tracks = [types.SimpleNamespace(observations=[0] * n_obs[i], covariance=np.eye(4, 4), prediction_at_frame=(lambda fr, i=i: pred[i].copy())) for i in range(T)]
        dets = [types.SimpleNamespace(to_xyah=(lambda j=j: det[j].copy())) for j in range(m)]
        kf = types.SimpleNamespace(gating_distance=lambda *a, **k: ns["gating_distance"](None, *a, **k))
        res = ns["fuse_motion"](kf, cost.copy(), tracks, dets, frame_id=7, use_lstm=True)
        out["fuse_motion"].append({"cost": cost, "pred": pred, "n_obs": n_obs, "det": det, "fused": np.asarray(res).copy()})
    torch.save(out, os.path.join(OUT, "tracker_logic.pt"))
    print("tracker logic ok:", {k: len(v["checks"]) for k, v in out["recorder"].items()}, len(out["fuse_motion"]), "fuse_motion cases")


def main():
    if not ref_loader.available():
        sys.exit("reference not present; goldens can only be generated in the build container")
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    gen_frame("frame_mot_96x128", spec.MOT_HEADS, "mot", 96, 128, N=10, n_prev=7, max_object=100, K=20)
    gen_frame("frame_kitti_64x160", spec.KITTI_HEADS, "kitti", 64, 160, N=6, n_prev=6, max_object=100, K=15)
    gen_frame("frame_nusc_64x96", spec.NUSC_HEADS, "nuscenes", 64, 96, N=9, n_prev=4, max_object=30, K=25)
    gen_dcn()
    gen_decode_plateau()
    gen_lstm()
    gen_afe_train()
    gen_tracker_logic()


if __name__ == "__main__":
    main()
