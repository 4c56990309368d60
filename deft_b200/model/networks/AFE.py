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
(deft_afe_affinity).  The training-side surface (forward/loss, AFE.py:74-86,235-328) is implemented as forward
values in eval mode (deft_afe_forward_scores / deft_afe_loss); gradients and train-mode BN are row f4.
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

    # -- a9 (training-side surface, forward pass only) ------------------------------------------
    def _extract_batched(self, sources, centres):
        """forward_selector_stacker1 (AFE.py:162-188) for a batch: sources = 13 maps [B,C,H,W], centres [B,N,1,1,2]
        (padding rows hold 1.5 -> border sample, image.py:356-359) -> [B,N,Cemb]."""
        pk = self.packed()
        maps = [m if isinstance(m, ops.NHWC) else ops.as_nhwc(m) for m in sources]
        B = maps[0].N
        N = centres.shape[1]
        out = torch.empty((B, N, sum(pk["sel_co"])), dtype=torch.float32, device=centres.device)
        ws = None
        for b in range(B):
            mb = [ops.NHWC(m.t[b:b + 1], m.C) for m in maps]
            ops.afe_extract(mb, pk["sel_w"], pk["sel_b"], pk["sel_co"], centres[b].reshape(-1, 2).float().contiguous(), self.align_corners,
                            out=out[b], ws=ws)
        return out

    def forward(self, sources_pre, sources_next, l_pre, l_next):
        """AFE.py:74-86: embeddings of both frames -> pairwise MLP over all P x P pairs -> [B,1,P+1,P+1] with the
        false-object column/row appended (add_unmatched_dim).  Eval-mode BN (running statistics): the train-mode
        batch statistics of ModleWithLoss (trainer.py:155-170) belong to the training engine."""
        _lib.require_device(l_pre)
        if self.training:
            raise NotImplementedError("AFE_module.forward: eval-mode BN only (call .eval()); train-mode batch statistics are row f4")
        P = self.max_object
        if l_pre.shape[1] != P or l_next.shape[1] != P:
            raise ValueError("AFE_module.forward: both centre tensors must hold max_object rows (generic_dataset pads them)")
        x_pre = self._extract_batched(sources_pre, l_pre)
        x_next = self._extract_batched(sources_next, l_next)
        B, _, Cemb = x_pre.shape
        pk = self.packed()["mlp"]
        au, _ = ops.afe_project(x_pre.view(B * P, Cemb), pk, want_bv=False, flags=self._flags, tcF=B, tcP=P)
        _, bv = ops.afe_project(x_next.view(B * P, Cemb), pk, want_au=False, flags=self._flags, tcF=B, tcP=P)
        return ops.afe_forward_scores(au.view(B, P, 512), bv.view(B, P, 512), pk, self.false_constant, flags=self._flags)

    def add_unmatched_dim(self, x):
        """AFE.py:215-233 on a device tensor [B,1,P,P] (plumbing; forward() already returns the extended matrix)."""
        ones_c = x.new_full(x.shape[:3] + (1,), self.false_constant)
        x = torch.cat([x, ones_c], 3)
        return torch.cat([x, x.new_full(x.shape[:2] + (1, x.shape[3]), self.false_constant)], 2)

    def loss(self, input, target, mask0, mask1):
        """AFE.py:235-328 (forward value): returns the reference's 8-tuple (loss_pre, loss_next, loss_similarity,
        mean loss, accuracy_pre, accuracy_next, mean accuracy, indexes_pre [B,1,P])."""
        _lib.require_device(input)
        if input.shape[-1] != self.max_object + 1:
            raise ValueError("AFE_module.loss: input must be [B,1,max_object+1,max_object+1]")
        dev = input.device
        u8 = lambda t: (t.to(dev) != 0).to(torch.uint8).contiguous()
        out7, idx = ops.afe_loss(input.float().contiguous(), target.to(dev).float().contiguous(), u8(mask0), u8(mask1))
        return (out7[0], out7[1], out7[2], out7[3], out7[4], out7[5], out7[6], idx)
