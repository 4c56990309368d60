"""One-time weight preparation for the C ABI (runs at model load / after load_state_dict, never per frame).

Layouts (see include/deft_b200.h):
  conv      W[Co,Ci,KH,KW]        -> Wp[(ky*KW+kx)*Ci+ci][ldw], ldw = ceil16(Co)
  eval BN   (w,b,mean,var,eps)    -> scale = w/sqrt(var+eps), shift = b - mean*scale
  DCN       weight/bias + actf BN -> Wp, scale, shift = bias*scale + bn_shift ; conv_offset_mask -> [9Ci][32]
  deconv    W[C,1,k,k]            -> Wp[ky*k+kx][C]
  heads     per-head (0.w,0.b,2.w,2.b) -> one [9*64][nh*256] + concatenated 1x1 rows
  AFE MLP   stacker2_bn and the three BNs folded into weights/shifts (AFE.py:190-213,331-347)
"""
from __future__ import annotations

import torch

BN_EPS = 1e-5


def ceil16(n):
    return (n + 15) // 16 * 16


def fold_bn(sd, p, eps=BN_EPS):
    scale = sd[p + ".weight"].float() / torch.sqrt(sd[p + ".running_var"].float() + eps)
    shift = sd[p + ".bias"].float() - sd[p + ".running_mean"].float() * scale
    return scale.contiguous(), shift.contiguous()


def pack_conv(w, ldw=None):
    co, ci, kh, kw = w.shape
    ldw = ceil16(co) if ldw is None else ldw
    wp = w.float().permute(2, 3, 1, 0).reshape(kh * kw * ci, co)
    if ldw != co:
        wp = torch.cat([wp, wp.new_zeros(wp.shape[0], ldw - co)], 1)
    return wp.contiguous(), ldw


def pad_vec(v, n):
    v = v.float()
    if v.numel() == n:
        return v.contiguous()
    return torch.cat([v, v.new_zeros(n - v.numel())]).contiguous()


def pack_dcn(sd, p, bn_prefix=None):
    """p = '<...>.conv' (the DCN); bn_prefix = '<...>.actf.0' or None for the bare op."""
    w = sd[p + ".weight"]
    wp, ldw = pack_conv(w)
    bias = sd[p + ".bias"].float()
    if bn_prefix is not None:
        scale, bshift = fold_bn(sd, bn_prefix)
        shift = bias * scale + bshift
    else:
        scale, shift = torch.ones_like(bias), bias.clone()
    w_om, _ = pack_conv(sd[p + ".conv_offset_mask.weight"], 32)
    b_om = pad_vec(sd[p + ".conv_offset_mask.bias"], 32)
    return dict(w=wp, ldw=ldw, scale=pad_vec(scale, ldw), shift=pad_vec(shift, ldw), w_om=w_om, b_om=b_om,
                cin=w.shape[1], cout=w.shape[0])


def pack_deconv(w):
    c, _, k, _ = w.shape
    return w.float().view(c, k * k).t().contiguous(), k // 2


def pack_heads(sd, heads):
    names = list(heads.keys())
    w1 = torch.cat([sd[f"{h}.0.weight"] for h in names], 0)
    w1p, ldw = pack_conv(w1)
    b1 = torch.cat([sd[f"{h}.0.bias"] for h in names]).float().contiguous()
    w2 = torch.cat([sd[f"{h}.2.weight"].float().view(heads[h], -1) for h in names], 0).contiguous()
    b2 = torch.cat([sd[f"{h}.2.bias"] for h in names]).float().contiguous()
    return dict(w1=w1p, b1=b1, w2=w2, b2=b2, head_c=[int(heads[h]) for h in names], names=names,
                head_conv=sd[f"{names[0]}.0.weight"].shape[0])


def pack_selector(sd, prefix="selector"):
    ws, bs, co = [], [], []
    for k in range(13):
        w = sd[f"{prefix}.{k}.weight"]
        wp, _ = pack_conv(w, w.shape[0])          # no column padding: [9*C][Co]
        ws.append(wp)
        bs.append(sd[f"{prefix}.{k}.bias"].float().contiguous())
        co.append(int(w.shape[0]))
    return ws, bs, co


def pack_affinity(sd, prefix=""):
    """Fold stacker2_bn (applied to both halves) and final_net BNs 1,4,7 (AFE.py:190-213,331-347)."""
    p = prefix
    s0, t0 = fold_bn(sd, p + "stacker2_bn")
    W1 = sd[p + "final_net.0.weight"].float().view(512, -1)
    b1 = sd[p + "final_net.0.bias"].float()
    cemb = W1.shape[1] // 2
    W1a, W1b = W1[:, :cemb], W1[:, cemb:]
    s1, t1 = fold_bn(sd, p + "final_net.1")
    w1a = (W1a * s0.view(1, -1) * s1.view(-1, 1)).t().contiguous()          # [Cemb][512]
    w1b = (W1b * s0.view(1, -1) * s1.view(-1, 1)).t().contiguous()
    sh1 = (s1 * (W1a @ t0 + W1b @ t0 + b1) + t1).contiguous()
    s2, t2 = fold_bn(sd, p + "final_net.4")
    w2 = sd[p + "final_net.3.weight"].float().view(256, 512).t().contiguous()
    sh2 = (s2 * sd[p + "final_net.3.bias"].float() + t2).contiguous()
    s3, t3 = fold_bn(sd, p + "final_net.7")
    w3 = sd[p + "final_net.6.weight"].float().view(128, 256).t().contiguous()
    sh3 = (s3 * sd[p + "final_net.6.bias"].float() + t3).contiguous()
    w4 = sd[p + "final_net.9.weight"].float().view(64, 128).t().contiguous()
    b4 = sd[p + "final_net.9.bias"].float().contiguous()
    w5 = sd[p + "final_net.11.weight"].float().view(64).contiguous()
    b5 = float(sd[p + "final_net.11.bias"].float().item())
    return dict(cemb=cemb, w1a=w1a, w1b=w1b, sh1=sh1, w2=w2, sc2=s2, sh2=sh2, w3=w3, sc3=s3, sh3=sh3, w4=w4, b4=b4, w5=w5, b5=b5)
