"""Drop-in for src/lib/model/networks/AFE.py ``AFE_module`` (the embedding extractor + pairwise-MLP
affinity of DEFT), B200-native.

Same constructor (``AFE_module(opt)``), same state_dict keys (``selector.{0..12}``, ``stacker2_bn``,
``final_net.{0,1,3,4,6,7,9,11}``), same call surface:
  forward_feature_extracter(FeatureMaps, centres[1,N,1,1,2]) -> [1,N,Cemb]     (AFE.py:88-92)
  forward_stacker_features(xp[1,n,C], xn[1,m,C], fill_up_column=True) -> np.float32[n, m+1(+n-1)] (AFE.py:110-160)
  get_similarity(...)                                                          (AFE.py:94-97)
plus the B200 batched form ``affinity_batched`` (all memory frames in one call, result stays on device).
The selector conv is evaluated only at the 4 bilinear corners of every centre (deft_afe_extract);
the MLP runs as chained GEMMs over the live n x m pairs with the first layer decomposed
(deft_afe_affinity).  Training-only pieces (forward/loss, AFE.py:74-86,235-328) are not on the
inference hot path and are not implemented here.
"""
import numpy as np
import torch
from torch import nn

from ... import _lib, ops, packing

SELECTOR_INPUT_CHANNEL = [16, 32, 64, 128, 256, 512, 64, 128, 256, 512, 64, 64, 64]


def _selector_out(dataset):
    if dataset == "nuscenes":
        return [48, 48, 64, 64, 64, 64, 64, 64, 64, 64, 32, 32, 32]
    return [32] * 13


class AFE_module(nn.Module):
    def __init__(self, opt):
        super().__init__()
        so = _selector_out(opt.dataset)
        emb = int(np.sum(so))
        widths = [2 * emb, 512, 256, 128, 64, 1]
        self.stacker2_bn = nn.BatchNorm2d(emb)
        self.final_dp = nn.Dropout(0.25)                     # parameter-free; kept for attribute parity
        layers = []
        for i in range(5):
            layers.append(nn.Conv2d(widths[i], widths[i + 1], kernel_size=1, stride=1))
            if i < 3:
                layers.append(nn.BatchNorm2d(widths[i + 1]))
            layers.append(nn.ReLU(inplace=True))
        self.final_net = nn.ModuleList(layers)               # conv at 0,3,6,9,11 ; BN at 1,4,7
        self.max_object = opt.max_object
        self.selector_input_channel = SELECTOR_INPUT_CHANNEL
        self.selector_output_channel = so
        self.selector = nn.ModuleList([nn.Conv2d(ci, co, kernel_size=3, padding=1) for ci, co in zip(SELECTOR_INPUT_CHANNEL, so)])
        self.false_constant = 1.0
        self.align_corners = bool(getattr(opt, "grid_sample_align_corners", False))
        self._flags = ops.FLAG_TENSORCORE if getattr(opt, "tensor_core", True) else 0
        if getattr(opt, "tf32x1", False):
            self._flags |= ops.FLAG_TF32X1
        self._pk = None
        self.register_load_state_dict_post_hook(lambda m, k: m._invalidate())

    # -- packed weights ---------------------------------------------------------------------
    def _invalidate(self):
        self._pk = None

    def _apply(self, fn, *a, **kw):
        self._pk = None
        return super()._apply(fn, *a, **kw)

    def packed(self):
        if self._pk is None:
            sd = {k: v.detach() for k, v in self.state_dict().items()}
            ws, bs, co = packing.pack_selector(sd)
            self._pk = dict(sel_w=ws, sel_b=bs, sel_co=co, mlp=packing.pack_affinity(sd))
        return self._pk

    # -- a7 ---------------------------------------------------------------------------------
    def forward_feature_extracter(self, s, l):
        """s: the 13 FeatureMaps ([1,C,H,W], any memory format); l: centres [1,N,1,1,2] in [-1,1]."""
        _lib.require_device(l)
        pk = self.packed()
        maps = [m if isinstance(m, ops.NHWC) else ops.as_nhwc(m) for m in s]
        if maps[0].N != 1:
            raise NotImplementedError("forward_feature_extracter: batch 1 (as in the reference's tracker)")
        centres = l.reshape(-1, 2).float().contiguous()
        emb = ops.afe_extract(maps, pk["sel_w"], pk["sel_b"], pk["sel_co"], centres, self.align_corners)
        return emb.unsqueeze(0)

    # -- a8 ---------------------------------------------------------------------------------
    def affinity_batched(self, mem, counts, cur, m=None, out=None, ws=None):
        """mem [F,P,Cemb] (device slab of F memory frames, rows >= counts[f] ignored), cur [>=m,Cemb].
        -> device tensor [F,P,P+1] (live block [:n_f, :m+1]); no host sync."""
        pk = self.packed()["mlp"]
        m = cur.shape[0] if m is None else m
        return ops.afe_affinity(mem, counts, cur, m, self.max_object, pk, out=out, ws=ws, flags=self._flags)

    def forward_stacker_features(self, xp, xn, fill_up_column=True):
        _lib.require_device(xp)
        P = self.max_object
        n, m = xp.shape[1], xn.shape[1]
        if n > P or m > P:
            raise ValueError("forward_stacker_features: n, m must be <= max_object (as in the reference)")
        mem = xp.reshape(1, n, -1).float()
        if n < P:
            mem = torch.cat([mem, mem.new_zeros(1, P - n, mem.shape[2])], 1)
        out = self.affinity_batched(mem.contiguous(), [n], xn.reshape(m, -1).float().contiguous(), m)
        x = out[0, :n, : m + 1]
        if fill_up_column and n > 1:
            x = torch.cat([x, x[:, m:m + 1].repeat(1, n - 1)], dim=1)
        return x.cpu().numpy()                                # the reference returns numpy (AFE.py:154-158)

    def get_similarity(self, image1, detection1, image2, detection2):
        f1 = self.forward_feature_extracter(image1, detection1)
        f2 = self.forward_feature_extracter(image2, detection2)
        return self.forward_stacker_features(f1, f2, False)

    def forward(self, *a, **k):
        raise NotImplementedError("AFE_module.forward/loss are training-only (AFE.py:74-86,235-328); out of the inference hot path")
