"""Drop-in for src/lib/model/networks/dla.py + base_model.py: ``DLASeg(num_layers, heads, head_convs, opt)``
with ``forward(x, pre_img=None, pre_hm=None) -> ([{head: Tensor}], FeatureMaps[13])`` and a state_dict
that is key-for-key / shape-for-shape the reference's (462 keys for the MOT heads; checkpoints
interchange in both directions).

The nn.Module tree below only HOLDS parameters under the reference's names.  The arithmetic is the
frame plan of deft_b200/engine.py: ~130 hand-written sm_100a kernel launches over an NHWC arena
(optionally one CUDA-graph replay).  There is no PyTorch-eager forward and no CPU path.

Returned tensors: heads are plain contiguous NCHW; the 13 FeatureMaps are [1,C,H,W]-shaped
channels_last views into the plan's arena (valid until the next forward of this model, like the
static outputs of a CUDA graph; set ``opt.clone_outputs`` for fresh copies).
"""
import numpy as np
import torch
from torch import nn

from ... import _lib, ops
from ...dcn_v2 import DCN
from ...engine import FramePlan
from .AFE import AFE_module

BN_MOMENTUM = 0.1


def _bn(c):
    return nn.BatchNorm2d(c, momentum=BN_MOMENTUM)


class BasicBlock(nn.Module):          # holder for dla.py:47-87
    def __init__(self, inplanes, planes, stride=1):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn1 = _bn(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=1, padding=1, bias=False)
        self.bn2 = _bn(planes)
        self.stride = stride


class Root(nn.Module):                # holder for dla.py:184-207
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, 1, bias=False)
        self.bn = _bn(out_channels)


class Tree(nn.Module):                # holder for dla.py:210-284
    def __init__(self, levels, in_channels, out_channels, stride=1, level_root=False, root_dim=0):
        super().__init__()
        if root_dim == 0:
            root_dim = 2 * out_channels
        if level_root:
            root_dim += in_channels
        if levels == 1:
            self.tree1 = BasicBlock(in_channels, out_channels, stride)
            self.tree2 = BasicBlock(out_channels, out_channels, 1)
            self.root = Root(root_dim, out_channels)
        else:
            self.tree1 = Tree(levels - 1, in_channels, out_channels, stride)
            self.tree2 = Tree(levels - 1, out_channels, out_channels, root_dim=root_dim + out_channels)
        if in_channels != out_channels:
            self.project = nn.Sequential(nn.Conv2d(in_channels, out_channels, 1, bias=False), _bn(out_channels))
        self.levels, self.level_root, self.root_dim = levels, level_root, root_dim


class DLA(nn.Module):                 # holder for dla.py:287-411 (dla34: levels [1,1,1,2,2,1])
    def __init__(self, levels, channels, opt=None):
        super().__init__()
        self.channels = channels
        c = channels
        self.base_layer = nn.Sequential(nn.Conv2d(3, c[0], 7, padding=3, bias=False), _bn(c[0]), nn.ReLU(inplace=True))
        self.level0 = nn.Sequential(nn.Conv2d(c[0], c[0], 3, padding=1, bias=False), _bn(c[0]), nn.ReLU(inplace=True))
        self.level1 = nn.Sequential(nn.Conv2d(c[0], c[1], 3, stride=2, padding=1, bias=False), _bn(c[1]), nn.ReLU(inplace=True))
        self.level2 = Tree(levels[2], c[1], c[2], 2, level_root=False)
        self.level3 = Tree(levels[3], c[2], c[3], 2, level_root=True)
        self.level4 = Tree(levels[4], c[3], c[4], 2, level_root=True)
        self.level5 = Tree(levels[5], c[4], c[5], 2, level_root=True)
        # real DEFT checkpoints carry these two (never used in any forward)
        if opt is not None and getattr(opt, "pre_img", False):
            self.pre_img_layer = nn.Sequential(nn.Conv2d(3, c[0], 7, padding=3, bias=False), _bn(c[0]), nn.ReLU(inplace=True))
        if opt is not None and getattr(opt, "pre_hm", False):
            self.pre_hm_layer = nn.Sequential(nn.Conv2d(1, c[0], 7, padding=3, bias=False), _bn(c[0]), nn.ReLU(inplace=True))


def dla34(pretrained=False, **kwargs):
    if pretrained:
        raise RuntimeError("deft_b200 never downloads ImageNet weights; pass a non-empty opt.load_model (dla.py:769-771)")
    return DLA([1, 1, 1, 2, 2, 1], [16, 32, 64, 128, 256, 512], **kwargs)


class DeformConv(nn.Module):          # dla.py:646-665
    def __init__(self, chi, cho):
        super().__init__()
        self.actf = nn.Sequential(_bn(cho), nn.ReLU(inplace=True))
        self.conv = DCN(chi, cho, kernel_size=(3, 3), stride=1, padding=1, dilation=1, deformable_groups=1)

    def forward(self, x):
        """Stand-alone use (the frame plan fuses BN+ReLU into the DCN epilogue the same way)."""
        from ... import packing
        _lib.require_device(x)
        sd = {"d." + k: v.detach() for k, v in self.state_dict().items()}
        pk = packing.pack_dcn(sd, "d.conv", "d.actf.0")
        return ops.dcnv2_forward(ops.as_nhwc(x), pk, relu=True).nchw_view()


def _fill_up_weights(up):
    w = up.weight.data
    k = w.size(2)
    f = int(np.ceil(k / 2))
    c = (2 * f - 1 - f % 2) / (2.0 * f)
    g = torch.tensor([1 - abs(i / f - c) for i in range(k)], dtype=w.dtype)
    w.copy_((g[:, None] * g[None, :]).expand_as(w))


class IDAUp(nn.Module):               # holder for dla.py:668-699
    def __init__(self, o, channels, up_f):
        super().__init__()
        for i in range(1, len(channels)):
            f = int(up_f[i])
            setattr(self, f"proj_{i}", DeformConv(channels[i], o))
            up = nn.ConvTranspose2d(o, o, f * 2, stride=f, padding=f // 2, output_padding=0, groups=o, bias=False)
            _fill_up_weights(up)
            setattr(self, f"up_{i}", up)
            setattr(self, f"node_{i}", DeformConv(o, o))


class DLAUp(nn.Module):               # holder for dla.py:702-735
    def __init__(self, startp, channels, scales):
        super().__init__()
        self.startp = startp
        in_channels = list(channels)
        channels = list(channels)
        scales = np.array(scales, dtype=int)
        for i in range(len(channels) - 1):
            j = -i - 2
            setattr(self, f"ida_{i}", IDAUp(channels[j], in_channels[j:], scales[j:] // scales[j]))
            scales[j + 1:] = scales[j]
            in_channels[j + 1:] = [channels[j] for _ in channels[j + 1:]]


class DLASeg(nn.Module):
    def __init__(self, num_layers, heads, head_convs, opt):
        super().__init__()
        if num_layers != 34:
            raise NotImplementedError("deft_b200 implements dla_34 (the only DLA DEFT's tracker works with)")
        if getattr(opt, "dla_node", "dcn") != "dcn":
            raise NotImplementedError("deft_b200 implements dla_node='dcn'")
        if getattr(opt, "head_kernel", 3) != 3:
            raise NotImplementedError("deft_b200 implements head_kernel=3")
        self.opt = opt
        self.heads = heads
        self.num_stacks = 1
        last_channel = 64
        for head in heads:                                       # base_model.py:33-103
            hc = head_convs[head]
            if len(hc) != 1 or hc[0] != 256:
                raise NotImplementedError("deft_b200 implements head_conv=[256] (DLA default, opts.py:449-450)")
            conv = nn.Conv2d(last_channel, hc[0], 3, padding=1, bias=True)
            out = nn.Conv2d(hc[0], heads[head], 1, bias=True)
            fc = nn.Sequential(conv, nn.ReLU(inplace=True), out)
            if "hm" in head:
                fc[-1].bias.data.fill_(opt.prior_bias)
            else:
                for m in fc.modules():
                    if isinstance(m, nn.Conv2d):
                        nn.init.constant_(m.bias, 0)
            setattr(self, head, fc)
        self.AFE = AFE_module(opt)
        self.first_level, self.last_level = 2, 5
        self.base = dla34(pretrained=(getattr(opt, "load_model", "x") == ""), opt=opt)
        channels = self.base.channels
        self.dla_up = DLAUp(self.first_level, channels[self.first_level:], [2 ** i for i in range(4)])
        self.ida_up = IDAUp(channels[self.first_level], channels[self.first_level:self.last_level], [2 ** i for i in range(3)])
        self._plans = {}
        self.use_cuda_graph = False
        self.register_load_state_dict_post_hook(lambda m, k: m._plans.clear())

    # -- plan management ----------------------------------------------------------------------
    def _apply(self, fn, *a, **kw):
        self._plans = {}
        return super()._apply(fn, *a, **kw)

    def invalidate(self):
        """Call after mutating parameters in place (load_state_dict / .to() do it automatically)."""
        self._plans = {}

    def _flags(self):
        fl = ops.FLAG_TENSORCORE if getattr(self.opt, "tensor_core", True) else 0
        if getattr(self.opt, "tf32x1", False):
            fl |= ops.FLAG_TF32X1
        return fl

    def plan_for(self, x, slot=0):
        """The frame plan (arena + launch list) for this input shape; `slot` > 0 asks for an additional, independent
        arena (FramePipeline keeps several so that frame t+1's backbone can overlap frame t's decode/embedding tail)."""
        return self.plan_for_shape(x.shape[0], x.shape[2], x.shape[3], x.device, slot)

    def plan_for_shape(self, N, H, W, device, slot=0, u8_norm=None):
        """u8_norm = (mean, std): a plan whose input slot is a uint8 HWC frame with the normalisation fused into the
        stem (row f2)."""
        # the PDL mode is part of the key: a captured graph bakes the launch attribute in (pipeline.py switches PDL off
        # when several frames are in flight)
        un = None if u8_norm is None else (tuple(float(v) for v in u8_norm[0]), tuple(float(v) for v in u8_norm[1]))
        key = ((N, 3, H, W), device.index, self._flags(), slot, int(_lib.lib().deft_get_pdl()), un)
        p = self._plans.get(key)
        if p is None:
            if self.training:
                raise _lib.DeftError("deft_b200.DLASeg runs eval-mode BN only: call .eval() (training is out of the hot path)")
            sd = {k: v for k, v in self.state_dict().items() if not k.startswith("AFE.")}
            with _lib.on_device(device):
                p = FramePlan(sd, self.heads, N, H, W, device, self._flags(), u8_norm=un)
                if self.use_cuda_graph:
                    p.capture()
            self._plans[key] = p
        return p

    # -- reference surface ----------------------------------------------------------------------
    def img2feats(self, x):
        """dla.py:789-802 -> ([y[-1]], FeatureMaps)."""
        _lib.require_device(x)
        p = self.plan_for(x)
        p.run(x)
        return [p.fm_views[-1]], list(p.fm_views)

    def imgpre2feats(self, x, pre_img=None, pre_hm=None):
        raise NotImplementedError("imgpre2feats is dead code in DEFT (Detector.pre_images is never set; SURVEY.md section 3.1)")

    @torch.no_grad()
    def forward(self, x, pre_img=None, pre_hm=None):
        if pre_img is not None or pre_hm is not None:
            return self.imgpre2feats(x, pre_img, pre_hm)
        _lib.require_device(x)
        if x.dtype != torch.float32 or x.dim() != 4 or x.shape[1] != 3:
            raise ValueError("DLASeg.forward expects a float32 [N,3,H,W] image batch")
        p = self.plan_for(x)
        heads, fms = p.run(x.contiguous())
        if getattr(self.opt, "clone_outputs", False):
            heads = {k: v.clone() for k, v in heads.items()}
            fms = [f.clone() for f in fms]
        if getattr(self.opt, "model_output_list", False):
            return [[heads[h] for h in sorted(self.heads)]], list(fms)
        return [dict((h, heads[h]) for h in self.heads)], list(fms)
