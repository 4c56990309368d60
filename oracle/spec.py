"""Parameter inventory + seeded synthetic initialiser (oracle side).

Names/shapes restate what the reference constructs in
  src/lib/model/networks/dla.py:287-361 (DLA), 210-284 (Tree), 184-207 (Root),
  47-87 (BasicBlock), 646-735 (DeformConv/IDAUp/DLAUp), 758-787 (DLASeg),
  src/lib/model/networks/base_model.py:24-103 (heads),
  src/lib/model/networks/AFE.py:18-72,331-364 (AFE_module),
  src/lib/utils/tracking_utils/kalman_filter_lstm.py:9-20 (DecoderRNN)
and the upstream ``dcn_v2.DCN`` parameter surface (weight, bias,
conv_offset_mask.{weight,bias}).  ``tests/golden/state_dict_keys_*.json`` holds
the key->shape table dumped from the *real* reference modules; the CPU test
suite checks this enumeration against it.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict

import torch

SELECTOR_IN = [16, 32, 64, 128, 256, 512, 64, 128, 256, 512, 64, 64, 64]  # AFE.py:15


def selector_out(dataset: str):
    # AFE.py:23-55
    if dataset == "nuscenes":
        return [48, 48, 64, 64, 64, 64, 64, 64, 64, 64, 32, 32, 32]
    return [32] * 13


MOT_HEADS = OrderedDict(hm=1, reg=2, wh=2, tracking=2, ltrb_amodal=4)            # opts.py:500-520
KITTI_HEADS = OrderedDict(hm=3, reg=2, wh=2, tracking=2)                         # kitti_tracking.sh:7
NUSC_HEADS = OrderedDict(hm=10, reg=2, wh=2, tracking=2, dep=1, rot=8, dim=3, amodel_offset=2)


def _bn(spec, p, c):
    spec[p + ".weight"] = ("bn_w", (c,))
    spec[p + ".bias"] = ("bn_b", (c,))
    spec[p + ".running_mean"] = ("bn_m", (c,))
    spec[p + ".running_var"] = ("bn_v", (c,))
    spec[p + ".num_batches_tracked"] = ("nbt", ())


def _block(spec, p, cin, cout):
    spec[p + ".conv1.weight"] = ("conv", (cout, cin, 3, 3))
    _bn(spec, p + ".bn1", cout)
    spec[p + ".conv2.weight"] = ("conv", (cout, cout, 3, 3))
    _bn(spec, p + ".bn2", cout)


def _tree1(spec, p, cin, cout, root_dim, project):
    """A levels==1 Tree: two BasicBlocks, a Root, optional project (dla.py:210-284)."""
    _block(spec, p + ".tree1", cin, cout)
    _block(spec, p + ".tree2", cout, cout)
    spec[p + ".root.conv.weight"] = ("conv", (cout, root_dim, 1, 1))
    _bn(spec, p + ".root.bn", cout)
    if project:
        spec[p + ".project.0.weight"] = ("conv", (cout, cin, 1, 1))
        _bn(spec, p + ".project.1", cout)


def _deform(spec, p, chi, cho):
    _bn(spec, p + ".actf.0", cho)
    spec[p + ".conv.weight"] = ("conv", (cho, chi, 3, 3))
    spec[p + ".conv.bias"] = ("bias", (cho,))
    spec[p + ".conv.conv_offset_mask.weight"] = ("om_w", (27, chi, 3, 3))
    spec[p + ".conv.conv_offset_mask.bias"] = ("om_b", (27,))


def _idaup(spec, p, o, channels, up_f):
    for i in range(1, len(channels)):
        f = int(up_f[i])
        _deform(spec, f"{p}.proj_{i}", channels[i], o)
        spec[f"{p}.up_{i}.weight"] = ("up", (o, 1, 2 * f, 2 * f))
        _deform(spec, f"{p}.node_{i}", o, o)


def dlaseg_spec(heads, head_conv=256, dataset="mot", pre_img=False, pre_hm=False):
    """key -> (kind, shape) for DLASeg(34, heads, {h:[head_conv]}, opt)."""
    spec = OrderedDict()
    for h, c in heads.items():                                   # base_model.py:33-103
        spec[f"{h}.0.weight"] = ("conv", (head_conv, 64, 3, 3))
        spec[f"{h}.0.bias"] = ("bias", (head_conv,))
        spec[f"{h}.2.weight"] = ("conv", (c, head_conv, 1, 1))
        spec[f"{h}.2.bias"] = ("hm_bias" if "hm" in h else "bias", (c,))
    so = selector_out(dataset)                                   # AFE.py
    emb = sum(so)
    _bn(spec, "AFE.stacker2_bn", emb)
    mlp = [2 * emb, 512, 256, 128, 64, 1]
    idx = 0
    for li in range(5):
        spec[f"AFE.final_net.{idx}.weight"] = ("conv", (mlp[li + 1], mlp[li], 1, 1))
        spec[f"AFE.final_net.{idx}.bias"] = ("bias", (mlp[li + 1],))
        if li < 3:
            _bn(spec, f"AFE.final_net.{idx + 1}", mlp[li + 1])
            idx += 3
        else:
            idx += 2
    for k in range(13):
        spec[f"AFE.selector.{k}.weight"] = ("conv", (so[k], SELECTOR_IN[k], 3, 3))
        spec[f"AFE.selector.{k}.bias"] = ("bias", (so[k],))
    ch = [16, 32, 64, 128, 256, 512]                             # dla.py:433-436
    spec["base.base_layer.0.weight"] = ("conv", (16, 3, 7, 7))
    _bn(spec, "base.base_layer.1", 16)
    spec["base.level0.0.weight"] = ("conv", (16, 16, 3, 3))
    _bn(spec, "base.level0.1", 16)
    spec["base.level1.0.weight"] = ("conv", (32, 16, 3, 3))
    _bn(spec, "base.level1.1", 32)
    _tree1(spec, "base.level2", 32, 64, 128, True)
    for lvl, cin, cout in ((3, 64, 128), (4, 128, 256)):
        p = f"base.level{lvl}"
        _tree1(spec, p + ".tree1", cin, cout, 2 * cout, True)
        _tree1(spec, p + ".tree2", cout, cout, 2 * cout + cin + cout, False)
        spec[p + ".project.0.weight"] = ("conv", (cout, cin, 1, 1))   # computed-but-unused in forward
        _bn(spec, p + ".project.1", cout)
    _tree1(spec, "base.level5", 256, 512, 1024 + 256, True)
    if pre_img:
        spec["base.pre_img_layer.0.weight"] = ("conv", (16, 3, 7, 7))
        _bn(spec, "base.pre_img_layer.1", 16)
    if pre_hm:
        spec["base.pre_hm_layer.0.weight"] = ("conv", (16, 1, 7, 7))
        _bn(spec, "base.pre_hm_layer.1", 16)
    # DLAUp(2, [64,128,256,512], [1,2,4,8])  dla.py:702-726
    _idaup(spec, "dla_up.ida_0", 256, [256, 512], [1, 2])
    _idaup(spec, "dla_up.ida_1", 128, [128, 256, 256], [1, 2, 2])
    _idaup(spec, "dla_up.ida_2", 64, [64, 128, 128, 128], [1, 2, 2, 2])
    _idaup(spec, "ida_up", 64, [64, 128, 256], [1, 2, 4])       # dla.py:782-787
    return spec


def lstm_spec(dataset="mot", hidden=128):
    fin, fout = (18, 16) if dataset == "nuscenes" else (11, 20)  # kalman_filter_lstm.py:13-20
    spec = OrderedDict()
    spec["lstm.weight_ih_l0"] = ("lin", (4 * hidden, fin))
    spec["lstm.weight_hh_l0"] = ("lin", (4 * hidden, hidden))
    spec["lstm.bias_ih_l0"] = ("lbias", (4 * hidden,))
    spec["lstm.bias_hh_l0"] = ("lbias", (4 * hidden,))
    spec["out1.weight"] = ("lin", (64, hidden))
    spec["out1.bias"] = ("lbias", (64,))
    spec["out2.weight"] = ("lin", (fout, 64))
    spec["out2.bias"] = ("lbias", (fout,))
    return spec


def _fill(kind, shape, g):
    """Synthetic distributions (with He-uniform conv gains so that
    activations keep O(1) spatial variation through ~60 stacked layers)."""
    if kind == "conv":
        fan_in = shape[1] * shape[2] * shape[3]
        b = math.sqrt(4.5 / fan_in)   # gain tuned so all 13 FeatureMaps keep std in [0.2, 4]
        return (torch.rand(shape, generator=g) * 2 - 1) * b
    if kind == "lin":
        b = 1.0 / math.sqrt(shape[1])
        return (torch.rand(shape, generator=g) * 2 - 1) * b
    if kind in ("bias", "bn_b", "bn_m", "lbias"):
        return torch.randn(shape, generator=g) * 0.1
    if kind == "hm_bias":
        return torch.full(shape, -4.6) + torch.randn(shape, generator=g) * 0.05
    if kind in ("bn_w", "bn_v"):
        return torch.rand(shape, generator=g) + 0.5
    if kind == "nbt":
        return torch.zeros((), dtype=torch.long)
    if kind == "om_w":
        return torch.randn(shape, generator=g) * 0.01
    if kind == "om_b":
        return torch.randn(shape, generator=g) * 0.75
    if kind == "up":
        # fill_up_weights (dla.py:565-573) bilinear kernel + a learnable-looking perturbation
        k = shape[2]
        f = math.ceil(k / 2)
        c = (2 * f - 1 - f % 2) / (2.0 * f)
        w1 = torch.tensor([1 - abs(i / f - c) for i in range(k)], dtype=torch.float32)
        w = (w1[:, None] * w1[None, :]).expand(shape).clone()
        return w + (torch.rand(shape, generator=g) * 2 - 1) * 0.05
    raise KeyError(kind)


def make_state_dict(spec, seed=317):
    """Deterministic (per key) fp32 state_dict.  Every tensor gets its own generator
    seeded from (seed, crc32(key)) so subsets/orderings never change the values."""
    sd = OrderedDict()
    for key, (kind, shape) in spec.items():
        g = torch.Generator()
        g.manual_seed((seed * 1000003 + zlib.crc32(key.encode())) & 0x7FFFFFFF)
        sd[key] = _fill(kind, tuple(shape), g)
    return sd


def make_input(shape, seed=317, tag="x"):
    g = torch.Generator()
    g.manual_seed((seed * 1000003 + zlib.crc32(tag.encode())) & 0x7FFFFFFF)
    return torch.randn(shape, generator=g)


def make_uniform(shape, lo, hi, seed=317, tag="u"):
    g = torch.Generator()
    g.manual_seed((seed * 1000003 + zlib.crc32(tag.encode())) & 0x7FFFFFFF)
    return torch.rand(shape, generator=g) * (hi - lo) + lo
