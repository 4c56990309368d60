"""Drop-in for the external ``dcn_v2`` module the reference imports at dla.py:25-29
(CharlesShang/DCNv2, un-vendored): ``DCN(chi, cho, kernel_size=(3,3), stride=1, padding=1, dilation=1,
deformable_groups=1)`` with parameters ``weight``, ``bias``, ``conv_offset_mask.{weight,bias}``.

forward() runs deft_dcnv2_forward (fused offset/mask conv + bilinear-gather GEMM) on the B200.
"""
import math

import torch
from torch import nn

from . import _lib, ops, packing


class _DCNFunction(torch.autograd.Function):
    """The DCN module under autograd (training, BASELINE config C5): forward = deft_dcnv2_forward (fp32 SIMT engine: exact
    arithmetic for gradient parity), backward = deft_dcnv2_backward (upstream dcn_v2_backward + the offset conv's backward)."""

    @staticmethod
    def forward(ctx, x, weight, bias, om_weight, om_bias):
        sd = {"d.weight": weight.detach(), "d.bias": bias.detach(), "d.conv_offset_mask.weight": om_weight.detach(),
              "d.conv_offset_mask.bias": om_bias.detach()}
        pk = packing.pack_dcn(sd, "d")
        a = ops.as_nhwc(x.detach())
        y = ops.dcnv2_forward(a, pk, relu=False, flags=0)
        ctx.pk, ctx.a = pk, a
        return y.nchw_view()

    @staticmethod
    def backward(ctx, grad_y):
        pk, a = ctx.pk, ctx.a
        cin, cout = pk["cin"], pk["cout"]
        gx, gw, gb, gwo, gbo = ops.dcnv2_backward(a, pk, ops.as_nhwc(grad_y.contiguous()))
        unpack = lambda g, co: g[:, :co].reshape(3, 3, cin, co).permute(3, 2, 0, 1).contiguous()      # [(ky,kx,ci)][co] -> [co,ci,ky,kx]
        return gx.nchw_view(), unpack(gw, cout), gb, unpack(gwo, 27), gbo[:27].contiguous()


class DCN(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=(3, 3), stride=1, padding=1, dilation=1, deformable_groups=1):
        super().__init__()
        ks = tuple(kernel_size) if isinstance(kernel_size, (tuple, list)) else (kernel_size, kernel_size)
        if ks != (3, 3) or stride != 1 or padding != 1 or dilation != 1 or deformable_groups != 1:
            raise NotImplementedError("deft_b200.DCN implements the configuration DEFT uses: 3x3, stride 1, pad 1, dil 1, 1 group")
        if in_channels % 16 != 0:
            raise NotImplementedError("deft_b200.DCN needs in_channels % 16 == 0")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels, 3, 3))
        self.bias = nn.Parameter(torch.zeros(out_channels))
        self.conv_offset_mask = nn.Conv2d(in_channels, 27, 3, stride=1, padding=1, bias=True)
        stdv = 1.0 / math.sqrt(in_channels * 9)                   # upstream reset_parameters / init_offset
        with torch.no_grad():
            self.weight.uniform_(-stdv, stdv)
            self.conv_offset_mask.weight.zero_()
            self.conv_offset_mask.bias.zero_()
        self._packed = None
        self.register_load_state_dict_post_hook(lambda m, k: m._invalidate())

    def _invalidate(self):
        self._packed = None

    def _apply(self, fn, *a, **kw):
        self._packed = None
        return super()._apply(fn, *a, **kw)

    def packed(self, bn_prefix_sd=None):
        if self._packed is None:
            sd = {"d." + k: v.detach() for k, v in self.state_dict().items()}
            self._packed = packing.pack_dcn(sd, "d")
        return self._packed

    def forward(self, x, flags=0):
        _lib.require_device(x)
        if torch.is_grad_enabled() and (x.requires_grad or self.weight.requires_grad):
            if self.out_channels % 16 != 0:
                raise NotImplementedError("deft_b200.DCN backward needs out_channels % 16 == 0")
            return _DCNFunction.apply(x, self.weight, self.bias, self.conv_offset_mask.weight, self.conv_offset_mask.bias)
        a = ops.as_nhwc(x)
        y = ops.dcnv2_forward(a, self.packed(), relu=False, flags=flags)
        return y.nchw_view()
