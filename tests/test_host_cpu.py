"""CPU suite, part 2: host logic without a GPU -- the C-ABI library loads and exports every symbol
include/deft_b200.h declares, the drop-in modules carry the reference's state_dict, weight packing is
algebraically right, and the product path fails loudly on CPU tensors (no fallback)."""
import json
import os
import re

import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN, ROOT, nerr
from oracle import restate, spec


def test_library_exports_every_declared_symbol():
    from deft_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "deft_b200.h")).read()
    declared = set(re.findall(r"\b(deft_[a-z0-9_]+)\s*\(", hdr))
    l = _lib.lib()                      # raises if the .so is missing
    for name in declared:
        assert hasattr(l, name), name
    assert declared == set(_lib.exported_symbols())
    assert l.deft_version() >= 100


def test_host_only_entry_points_without_a_gpu(monkeypatch):
    """The C-ABI calls that never touch the device: launch policy switch, tile / blob / scratch sizing."""
    from deft_b200 import _lib
    l = _lib.lib()
    if "DEFT_PDL" not in os.environ:
        l.deft_set_pdl(0)
        assert l.deft_get_pdl() == 0
        l.deft_set_pdl(1)
        assert l.deft_get_pdl() == 1
    M = 152 * 272
    for ldw, kind in ((64, 0), (128, 0), (32, 0), (64, 1)):
        bn = l.deft_tc_tile_n(ldw, M, kind)
        assert bn in (16, 32, 64, 128) and ldw % bn == 0
        assert l.deft_tc_blob_bytes(ldw, 9 * 64, bn) == (ldw // bn) * (9 * 64 // 32) * 2 * bn * 32 * 4     # hi + lo tile images
    assert l.deft_tc_tile_n(1280, M, 2) == 256                                                         # fused heads: one head per N tile
    assert l.deft_dcnv2_workspace_bytes(1, 152, 272, 64, 64) >= M * 32 * 4                               # at least the offset/mask map
    assert l.deft_tc_scratch_bytes(512, 9 * 512, 19 * 34, 0) >= 2 * 19 * 34 * 512 * 4                    # deep layer: split-K partials
    # split-K policy (tc_pick_splits): a DCN gather-GEMM of the same shape is never given FEWER partial buffers than a conv
    # (DCN layers split up to half the SMs, convs up to a third) and a map with >= 100 tiles gets none at all
    assert l.deft_tc_scratch_bytes(128, 9 * 256, 38 * 68, 1) >= l.deft_tc_scratch_bytes(128, 9 * 256, 38 * 68, 0) > 0
    assert l.deft_tc_scratch_bytes(64, 9 * 64, M, 0) == 0 and l.deft_tc_scratch_bytes(64, 9 * 64, M, 1) == 0
    # the second DCN gather kernel is opt-in and its launch counter starts at zero in a process that never launched it
    assert l.deft_dcn_patch_launch_count() == 0
    l.deft_debug_set_dcn_patch(2)
    l.deft_debug_set_dcn_patch(0)


@pytest.mark.parametrize("heads,dataset", [(spec.MOT_HEADS, "mot"), (spec.KITTI_HEADS, "kitti"), (spec.NUSC_HEADS, "nuscenes")])
def test_dropin_state_dict_matches_reference(heads, dataset):
    from deft_b200.model.model import create_model
    from deft_b200.opts import Opt
    m = create_model("dla_34", dict(heads), {h: [256] for h in heads}, Opt(dataset=dataset, heads=dict(heads)))
    with open(os.path.join(GOLDEN, f"state_dict_keys_{dataset}_{len(heads)}heads.json")) as f:
        ref = {k: tuple(v) for k, v in json.load(f).items()}
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == ref
    m.load_state_dict(spec.make_state_dict(spec.dlaseg_spec(heads, dataset=dataset)), strict=True)


def test_checkpoint_with_pre_img_layers_and_module_prefix(tmp_path):
    """Real DEFT checkpoints carry base.pre_img_layer/pre_hm_layer and may be 'module.'-prefixed (model.py:48-52)."""
    from deft_b200.model.model import create_model, load_model, save_model
    from deft_b200.opts import Opt
    heads = spec.MOT_HEADS
    sd = spec.make_state_dict(spec.dlaseg_spec(heads, pre_img=True, pre_hm=True))
    path = str(tmp_path / "ckpt.pth")
    torch.save({"epoch": 7, "state_dict": {"module." + k: v for k, v in sd.items()}}, path)
    opt = Opt(pre_img=True, pre_hm=True)
    m = create_model("dla_34", dict(heads), {h: [256] for h in heads}, opt)
    m = load_model(m, path, opt)
    got = m.state_dict()
    assert set(got) == set(sd)
    assert all(torch.equal(got[k], sd[k]) for k in sd)
    save_model(str(tmp_path / "out.pth"), 8, m)
    assert set(torch.load(str(tmp_path / "out.pth"), weights_only=False)["state_dict"]) == set(sd)
    m2 = create_model("dla_34", dict(heads), {h: [256] for h in heads}, Opt())   # without the dead layers: they are dropped
    load_model(m2, path, Opt())


def test_load_model_resume_semantics_and_shape_mismatch(tmp_path, capsys):
    """model.py:56-110: a mis-shaped tensor is reported and skipped; with an optimizer the answer is always
    (model, optimizer, start_epoch), start_epoch 0 unless opt.resume finds an optimizer entry, and the learning rate
    is re-derived from opt.lr / opt.lr_step."""
    import types
    from deft_b200.model.model import create_model, load_model, save_model
    from deft_b200.opts import Opt
    heads = spec.MOT_HEADS
    m = create_model("dla_34", dict(heads), {h: [256] for h in heads}, Opt())
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    bad = "hm.2.weight"
    sd[bad] = torch.zeros(3, 256, 1, 1)                       # wrong number of classes
    sd["not.a.key"] = torch.zeros(1)
    opt_sgd = torch.optim.SGD(m.parameters(), lr=1.0)
    path = str(tmp_path / "c.pth")
    torch.save({"epoch": 65, "state_dict": sd, "optimizer": opt_sgd.state_dict()}, path)
    keep = m.state_dict()[bad].clone()
    out = load_model(m, path, types.SimpleNamespace(resume=False, lr=1.25e-4, lr_step=[60, 70]), optimizer=opt_sgd)
    assert isinstance(out, tuple) and out[2] == 0 and opt_sgd.param_groups[0]["lr"] == 1.0
    assert torch.equal(m.state_dict()[bad], keep)
    log = capsys.readouterr().out
    assert "Skip loading parameter " + bad in log and "Drop parameter not.a.key" in log
    out = load_model(m, path, types.SimpleNamespace(resume=True, lr=1.25e-4, lr_step=[60, 70]), optimizer=opt_sgd)
    assert out[2] == 65 and abs(opt_sgd.param_groups[0]["lr"] - 1.25e-5) < 1e-12
    save_model(str(tmp_path / "d.pth"), 66, torch.nn.DataParallel(m), opt_sgd)
    d = torch.load(str(tmp_path / "d.pth"), weights_only=False)
    assert d["epoch"] == 66 and "optimizer" in d and not any(k.startswith("module.") for k in d["state_dict"])


def test_no_cpu_fallback():
    from deft_b200._lib import DeftError
    from deft_b200.kalman_filter_lstm import DecoderRNN
    from deft_b200.model.decode import generic_decode
    from deft_b200.model.model import create_model
    from deft_b200.opts import Opt
    heads = spec.MOT_HEADS
    m = create_model("dla_34", dict(heads), {h: [256] for h in heads}, Opt()).eval()
    with pytest.raises(DeftError):
        m(torch.zeros(1, 3, 64, 64))
    with pytest.raises(DeftError):
        m.AFE.forward_feature_extracter([torch.zeros(1, 16, 8, 8)] * 13, torch.zeros(1, 2, 1, 1, 2))
    with pytest.raises(DeftError):
        m.AFE.forward_stacker_features(torch.zeros(1, 2, 416), torch.zeros(1, 2, 416))
    with pytest.raises(DeftError):
        generic_decode({"hm": torch.zeros(1, 1, 8, 8)}, K=4)
    with pytest.raises(DeftError):
        DecoderRNN(128, "mot").step(torch.zeros(1, 11), torch.zeros(1, 128), torch.zeros(1, 128))
    with pytest.raises(NotImplementedError):
        create_model("resdcn_18", dict(heads), {h: [256] for h in heads}, Opt())


def test_product_never_imports_oracle():
    for dp, _, fs in os.walk(os.path.join(ROOT, "deft_b200")):
        for f in fs:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+\.*oracle", src, re.M), os.path.join(dp, f)
                assert "import oracle" not in src and "from oracle" not in src, os.path.join(dp, f)


def test_packing_algebra():
    """pack_conv / fold_bn / pack_affinity reproduce the unfolded arithmetic (CPU, fp32)."""
    from deft_b200 import packing
    g = torch.Generator(); g.manual_seed(1)
    w = torch.randn(24, 16, 3, 3, generator=g)
    wp, ldw = packing.pack_conv(w)
    assert ldw == 32 and wp.shape == (144, 32) and torch.equal(wp[:, 24:], torch.zeros(144, 8))
    x = torch.randn(1, 16, 5, 6, generator=g)
    cols = F.unfold(x, 3, padding=1).view(16, 9, 30).permute(1, 0, 2).reshape(144, 30)     # k = tap*Cin + ci
    assert nerr(wp[:, :24].t() @ cols, F.conv2d(x, w, padding=1).view(24, 30)) < 1e-5
    sd = spec.make_state_dict({k: v for k, v in spec.dlaseg_spec(spec.MOT_HEADS).items() if k.startswith("AFE.")})
    pk = packing.pack_affinity({k[4:]: v for k, v in sd.items()})
    xp, xn = spec.make_input((1, 4, 416), tag="pa").abs(), spec.make_input((1, 3, 416), tag="pb").abs()
    au = xp[0] @ pk["w1a"] + pk["sh1"]
    bv = xn[0] @ pk["w1b"]
    h = torch.relu(au[:, None, :] + bv[None, :, :]).reshape(12, 512)
    h = torch.relu((h @ pk["w2"]) * pk["sc2"] + pk["sh2"])
    h = torch.relu((h @ pk["w3"]) * pk["sc3"] + pk["sh3"])
    h = torch.relu(h @ pk["w4"] + pk["b4"])
    S = torch.relu(h @ pk["w5"] + pk["b5"]).view(4, 3)
    ref = restate.afe_affinity(sd, xp, xn, 100, False)
    X = torch.zeros(101, 101); X[:4, :3] = S; X[100, :] = 1; X[:, 100] = 1
    mine = torch.max(torch.softmax(X, 1), torch.softmax(X, 0))[:4, :3]
    assert nerr(mine, ref[:, :3]) < 1e-5


def test_load_model_reset_hm_and_reuse_hm(tmp_path, capsys):
    """model.py:58-77: with opt.reset_hm an `hm*` tensor whose leading dimension is 80 (COCO) or 1 (CrowdHuman) is treated as a
    mismatch even when the shapes agree (-> skipped); with opt.reuse_hm a mismatched tensor's leading slice is reused."""
    import torch
    from types import SimpleNamespace
    from deft_b200.model.model import _reconcile
    own = {"hm.2.weight": torch.zeros(1, 256, 1, 1), "hm.2.bias": torch.zeros(1), "wh.2.bias": torch.zeros(2)}
    ck = {"hm.2.weight": torch.ones(1, 256, 1, 1), "hm.2.bias": torch.ones(1), "wh.2.bias": torch.ones(2)}
    merged, rep = _reconcile(ck, own)
    assert merged["hm.2.bias"].item() == 1 and not rep
    merged, rep = _reconcile(ck, own, reset_hm=True)
    assert merged["hm.2.bias"].item() == 0 and merged["hm.2.weight"].sum().item() == 0 and merged["wh.2.bias"].sum().item() == 2
    assert sum("Skip loading parameter hm" in r for r in rep) == 2
    coco = {"hm.2.weight": torch.arange(80.).view(80, 1, 1, 1).repeat(1, 256, 1, 1), "hm.2.bias": torch.arange(80.)}
    merged, rep = _reconcile(coco, own, reuse_hm=True)
    assert merged["hm.2.bias"].tolist() == [0.0] and merged["hm.2.weight"].shape == (1, 256, 1, 1)
    assert all("Reusing parameter" in r for r in rep[:2])


def test_bench_stream_assignment_for_the_nuscenes_split():
    """BASELINE config C4: 6 cameras dealt round-robin over 2/4/8 GPUs; every stream is owned exactly once."""
    from deft_b200.parallel import stream_assignment
    for world in (1, 2, 4, 8):
        owned = [stream_assignment(6, world, r) for r in range(world)]
        assert sorted(sum(owned, [])) == list(range(6))
        assert max(len(o) for o in owned) == -(-6 // world)
    import bench
    nf, ni, nc = bench.synthetic_tracks(500, 5, 100, 6, 5)
    assert int(nc.min()) == 1 and int(nc.max()) == 5 and int(nf[nf >= 0].min()) == 0 and int(nf.max()) == 4


def test_bench_cpu_arm_picks_a_usable_thread_count():
    """bench.py's CPU arm (cpu_baseline / --impl reference) must run the reference at its best on the box: never more threads
    than the affinity mask / cgroup quota allow, and the probe keeps the fastest of the candidate counts (a 128-CPU container
    with 128 threads was 12x slower per frame than 8 threads here)."""
    import bench
    n = bench.usable_cores()
    assert 1 <= n <= (os.cpu_count() or 1) and n <= len(os.sched_getaffinity(0))
    before = torch.get_num_threads()
    try:
        best, table = bench.pick_cpu_threads(torch)
        assert best in table and max(table) == n and torch.get_num_threads() == best
        assert all(v > 0 for v in table.values())
        assert table[best] <= min(table.values()) / 0.97 + 1e-9          # within the 3 % preference for more threads
    finally:
        torch.set_num_threads(before)
