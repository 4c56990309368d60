"""Stand-in for the un-vendored CharlesShang/DCNv2 ``dcn_v2`` module (reference README.md:72-78),
used ONLY by oracle/ref_loader.py so the reference's own DLASeg can be constructed on CPU.

Same parameter surface as upstream ``DCN`` (weight, bias, conv_offset_mask.{weight,bias}); the
arithmetic is torchvision.ops.deform_conv2d (same offset/mask convention).
TEST INFRASTRUCTURE ONLY.
"""
import math

import torch
from torch import nn
from torchvision.ops import deform_conv2d


class DCN(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=(3, 3), stride=1, padding=1,
                 dilation=1, deformable_groups=1):
        super().__init__()
        kh, kw = kernel_size
        self.stride, self.padding, self.dilation = stride, padding, dilation
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels, kh, kw))
        self.bias = nn.Parameter(torch.zeros(out_channels))
        self.conv_offset_mask = nn.Conv2d(in_channels, deformable_groups * 3 * kh * kw, (kh, kw),
                                          stride=stride, padding=padding, bias=True)
        stdv = 1.0 / math.sqrt(in_channels * kh * kw)
        self.weight.data.uniform_(-stdv, stdv)
        self.conv_offset_mask.weight.data.zero_()
        self.conv_offset_mask.bias.data.zero_()

    def forward(self, x):
        out = self.conv_offset_mask(x)
        o1, o2, mask = torch.chunk(out, 3, dim=1)
        offset = torch.cat((o1, o2), dim=1)
        mask = torch.sigmoid(mask)
        return deform_conv2d(x, offset, self.weight, self.bias, stride=self.stride,
                             padding=self.padding, dilation=self.dilation, mask=mask)
