"""fp32 PyTorch-CPU restatement of the DEFT per-frame dense path, functional over a state_dict.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Every function cites the
reference file:line (relative to /root/reference/src/lib) whose arithmetic it restates.
It is pinned against the real reference modules by oracle/gen_golden.py and
tests/test_oracle_cpu.py (tests/golden/*.pt).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

EPS = 1e-5  # nn.BatchNorm2d default eps


# --------------------------------------------------------------------------- primitives
def bn_eval(sd, p, x):
    """Eval-mode BatchNorm2d: (x-mean)/sqrt(var+eps)*w+b  (torch.nn.BatchNorm2d, used at
    dla.py:63,74,196,...)."""
    sh = (1, -1, 1, 1)
    return (x - sd[p + ".running_mean"].view(sh)) / torch.sqrt(sd[p + ".running_var"].view(sh) + EPS) \
        * sd[p + ".weight"].view(sh) + sd[p + ".bias"].view(sh)


def dcn_v2_explicit(x, offset, mask, weight, bias):
    """Modulated deformable conv v2, 3x3/stride1/pad1/dil1/1 deformable group, written out.

    Upstream CharlesShang/DCNv2 ``dcn_v2_forward`` semantics (un-vendored; call site
    dla.py:652-663):  y[co,h,w] = b[co] + sum_{ci,k} W[co,ci,k] * m[k,h,w] *
    bilinear(x[ci]; h-1+ky+off[2k], w-1+kx+off[2k+1]) with k = ky*3+kx, the sample being
    zero when the point is outside (-1,H)x(-1,W) and each of the 4 corners contributing
    only if it lies inside the image.
    """
    B, C, H, W = x.shape
    ys = torch.arange(H, dtype=torch.float32).view(1, H, 1)
    xs = torch.arange(W, dtype=torch.float32).view(1, 1, W)
    cols = []
    xf = x.reshape(B, C, H * W)
    for k in range(9):
        ky, kx = divmod(k, 3)
        py = ys - 1 + ky + offset[:, 2 * k]
        px = xs - 1 + kx + offset[:, 2 * k + 1]
        y0 = torch.floor(py)
        x0 = torch.floor(px)
        ly, lx = py - y0, px - x0
        val = torch.zeros(B, C, H, W)
        inside = (py > -1) & (py < H) & (px > -1) & (px < W)
        for dy, wy in ((0, 1 - ly), (1, ly)):
            for dx, wx in ((0, 1 - lx), (1, lx)):
                yy = (y0 + dy).long()
                xx = (x0 + dx).long()
                ok = inside & (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
                idx = (yy.clamp(0, H - 1) * W + xx.clamp(0, W - 1)).view(B, 1, H * W).expand(B, C, H * W)
                g = torch.gather(xf, 2, idx).view(B, C, H, W)
                val = val + g * (wy * wx * ok.float()).unsqueeze(1)
        cols.append(val * mask[:, k].unsqueeze(1))
    col = torch.stack(cols, 2)                                  # [B,C,9,H,W]
    out = torch.einsum("ock,bckhw->bohw", weight.reshape(weight.shape[0], C, 9), col)
    return out + bias.view(1, -1, 1, 1)


def dcn_forward(sd, p, x, explicit=False):
    """dcn_v2.DCN.forward (upstream): om = conv_offset_mask(x); o1,o2,m = chunk(om,3,1);
    offset = cat(o1,o2); mask = sigmoid(m); modulated deformable conv."""
    om = F.conv2d(x, sd[p + ".conv_offset_mask.weight"], sd[p + ".conv_offset_mask.bias"], padding=1)
    o1, o2, m = torch.chunk(om, 3, dim=1)
    offset = torch.cat((o1, o2), dim=1)
    mask = torch.sigmoid(m)
    if explicit:
        return dcn_v2_explicit(x, offset, mask, sd[p + ".weight"], sd[p + ".bias"])
    from torchvision.ops import deform_conv2d
    return deform_conv2d(x, offset, sd[p + ".weight"], sd[p + ".bias"], stride=1, padding=1,
                         dilation=1, mask=mask)


def deform_conv(sd, p, x, explicit=False):
    """DeformConv.forward = DCN -> BN -> ReLU   (dla.py:646-665)."""
    return F.relu(bn_eval(sd, p + ".actf.0", dcn_forward(sd, p + ".conv", x, explicit)))


# --------------------------------------------------------------------------- backbone
def basic_block(sd, p, x, residual=None, stride=1):
    """BasicBlock.forward (dla.py:73-87)."""
    if residual is None:
        residual = x
    out = F.relu(bn_eval(sd, p + ".bn1", F.conv2d(x, sd[p + ".conv1.weight"], stride=stride, padding=1)))
    out = bn_eval(sd, p + ".bn2", F.conv2d(out, sd[p + ".conv2.weight"], padding=1))
    return F.relu(out + residual)


def root(sd, p, *xs):
    """Root.forward with residual=False for dla34 (dla.py:199-207)."""
    return F.relu(bn_eval(sd, p + ".bn", F.conv2d(torch.cat(xs, 1), sd[p + ".conv.weight"])))


def tree(sd, p, levels, x, stride, level_root, children=None):
    """Tree.forward (dla.py:271-284).  Note the ``residual`` argument the reference passes
    down is recomputed (and so ignored) by every callee."""
    children = [] if children is None else children
    bottom = F.max_pool2d(x, stride, stride) if stride > 1 else x
    has_proj = (p + ".project.0.weight") in sd
    if level_root:
        children.append(bottom)
    if levels == 1:
        residual = bn_eval(sd, p + ".project.1", F.conv2d(bottom, sd[p + ".project.0.weight"])) if has_proj else bottom
        x1 = basic_block(sd, p + ".tree1", x, residual, stride)
        x2 = basic_block(sd, p + ".tree2", x1)
        return root(sd, p + ".root", x2, x1, *children)
    x1 = tree(sd, p + ".tree1", levels - 1, x, stride, False)
    children.append(x1)
    return tree(sd, p + ".tree2", levels - 1, x1, 1, False, children=children)


def dla34_base(sd, x, p="base"):
    """DLA.forward for dla34 levels [1,1,1,2,2,1] (dla.py:400-411, 433-441)."""
    y = []
    x = F.relu(bn_eval(sd, p + ".base_layer.1", F.conv2d(x, sd[p + ".base_layer.0.weight"], padding=3)))
    x = F.relu(bn_eval(sd, p + ".level0.1", F.conv2d(x, sd[p + ".level0.0.weight"], padding=1)))
    y.append(x)
    x = F.relu(bn_eval(sd, p + ".level1.1", F.conv2d(x, sd[p + ".level1.0.weight"], stride=2, padding=1)))
    y.append(x)
    x = tree(sd, p + ".level2", 1, x, 2, False); y.append(x)
    x = tree(sd, p + ".level3", 2, x, 2, True); y.append(x)
    x = tree(sd, p + ".level4", 2, x, 2, True); y.append(x)
    x = tree(sd, p + ".level5", 1, x, 2, True); y.append(x)
    return y


def up_deconv(sd, p, x):
    """Depthwise ConvTranspose2d(o,o,2f,stride=f,padding=f//2,groups=o,bias=False) (dla.py:677-687)."""
    w = sd[p + ".weight"]
    f = w.shape[2] // 2
    return F.conv_transpose2d(x, w, stride=f, padding=f // 2, groups=w.shape[0])


def ida_up(sd, p, layers, startp, endp, explicit=False):
    """IDAUp.forward (dla.py:693-699), mutating ``layers`` like the reference."""
    for i in range(startp + 1, endp):
        j = i - startp
        layers[i] = up_deconv(sd, f"{p}.up_{j}", deform_conv(sd, f"{p}.proj_{j}", layers[i], explicit))
        layers[i] = deform_conv(sd, f"{p}.node_{j}", layers[i] + layers[i - 1], explicit)


def dla_up(sd, layers, explicit=False, p="dla_up", startp=2):
    """DLAUp.forward (dla.py:728-735)."""
    layers = list(layers)
    out = [layers[-1]]
    for i in range(len(layers) - startp - 1):
        ida_up(sd, f"{p}.ida_{i}", layers, len(layers) - i - 2, len(layers), explicit)
        out.insert(0, layers[-1])
    return out


def img2feats(sd, x, explicit=False):
    """DLASeg.img2feats (dla.py:789-802): returns ([y[-1]], 13 FeatureMaps)."""
    base = dla34_base(sd, x)
    fm = list(base)
    ups = dla_up(sd, base, explicit)
    fm += ups
    y = [ups[i].clone() for i in range(3)]
    ida_up(sd, "ida_up", y, 0, 3, explicit)
    fm += y
    return [y[-1]], fm


def head_forward(sd, h, feat):
    """One head: conv3x3(64->256)+bias -> ReLU -> conv1x1(256->C)+bias (base_model.py:33-68)."""
    t = F.relu(F.conv2d(feat, sd[f"{h}.0.weight"], sd[f"{h}.0.bias"], padding=1))
    return F.conv2d(t, sd[f"{h}.2.weight"], sd[f"{h}.2.bias"])


def dlaseg_forward(sd, x, heads, explicit=False):
    """BaseModel.forward with pre_img=pre_hm=None, model_output_list=False
    (base_model.py:111-132) -> ([{head: Tensor}], FeatureMaps[13])."""
    feats, fm = img2feats(sd, x, explicit)
    return [{h: head_forward(sd, h, feats[0]) for h in heads}], fm


# --------------------------------------------------------------------------- decode
def sigmoid_output(output, depth_scale=1.0):
    """Detector._sigmoid_output (detector.py:486-494)."""
    out = dict(output)
    if "hm" in out:
        out["hm"] = torch.sigmoid(out["hm"])
    if "dep" in out:
        out["dep"] = (1.0 / (torch.sigmoid(out["dep"]) + 1e-6) - 1.0) * depth_scale
    return out


def _gather(feat, ind):
    """_tranpose_and_gather_feat (model/utils.py:25-36)."""
    B, C = feat.shape[:2]
    f = feat.permute(0, 2, 3, 1).reshape(B, -1, C)
    return f.gather(1, ind.unsqueeze(2).expand(B, ind.shape[1], C))


def generic_decode(output, K=100, zero_tracking=False):
    """generic_decode + _nms + _topk (decode.py:102-219, utils.py:69-104), restricted to the
    heads DEFT's three datasets use (no hps).  Tie rule among equal scores: lowest flat index."""
    heat = output["hm"]
    B, C, H, W = heat.shape
    hmax = F.max_pool2d(heat, 3, stride=1, padding=1)
    heat = heat * (hmax == heat).float()
    flat = heat.view(B, C, -1)
    # stable descending sort == "lowest flat index wins" for equal values
    s1, i1 = torch.sort(flat, dim=2, descending=True, stable=True)
    s1, i1 = s1[:, :, :K], i1[:, :, :K]
    s2, i2 = torch.sort(s1.reshape(B, -1), dim=1, descending=True, stable=True)
    scores, i2 = s2[:, :K], i2[:, :K]
    clses = torch.div(i2, K, rounding_mode="floor").int()
    inds = i1.reshape(B, -1).gather(1, i2)
    ys0 = torch.div(inds, W, rounding_mode="floor").float()
    xs0 = (inds % W).float()
    ret = {"scores": scores, "clses": clses.float(), "xs": xs0, "ys": ys0,
           "cts": torch.stack([xs0, ys0], dim=2), "inds": inds}
    if "reg" in output:
        reg = _gather(output["reg"], inds)
        xs = xs0.unsqueeze(2) + reg[:, :, 0:1]
        ys = ys0.unsqueeze(2) + reg[:, :, 1:2]
    else:
        xs = xs0.unsqueeze(2) + 0.5
        ys = ys0.unsqueeze(2) + 0.5
    if "wh" in output:
        wh = _gather(output["wh"], inds).clone()
        wh[wh < 0] = 0
        ret["bboxes"] = torch.cat([xs - wh[..., 0:1] / 2, ys - wh[..., 1:2] / 2,
                                   xs + wh[..., 0:1] / 2, ys + wh[..., 1:2] / 2], dim=2)
    for h in ("tracking", "dep", "rot", "dim", "amodel_offset"):
        if h in output:
            v = _gather(output[h], inds)
            ret[h] = v * 0 if (h == "tracking" and zero_tracking) else v
    if "ltrb_amodal" in output:
        l = _gather(output["ltrb_amodal"], inds)
        bb = torch.cat([xs0.unsqueeze(2) + l[..., 0:1], ys0.unsqueeze(2) + l[..., 1:2],
                        xs0.unsqueeze(2) + l[..., 2:3], ys0.unsqueeze(2) + l[..., 3:4]], dim=2)
        ret["bboxes_amodal"] = bb
        ret["bboxes"] = bb
    return ret


# --------------------------------------------------------------------------- AFE
def afe_extract(sd, feature_maps, centres, align_corners=False):
    """AFE_module.forward_feature_extracter -> forward_selector_stacker1 (AFE.py:88-92,162-188):
    13 x relu(conv3x3) then a bilinear border-padded sample per detection, concatenated.
    centres: [1,N,1,1,2] (x,y) in [-1,1].  Returns [1,N,C_emb]."""
    outs = []
    grid = centres[:, :, 0, :, :]                                # [B,N,1,2] -> all N points at once
    for k, fmap in enumerate(feature_maps):
        s = F.relu(F.conv2d(fmap, sd[f"AFE.selector.{k}.weight"], sd[f"AFE.selector.{k}.bias"], padding=1))
        g = F.grid_sample(s, grid, mode="bilinear", padding_mode="border", align_corners=align_corners)
        outs.append(g[:, :, :, 0].permute(0, 2, 1))             # [B,N,C_k]
    return torch.cat(outs, 2)


def afe_affinity(sd, xp, xn, max_object, fill_up_column=True):
    """AFE_module.forward_stacker_features (AFE.py:110-160) with forward_stacker2 (190-207),
    forward_final (209-213), add_final (331-347).  xp [1,n,C], xn [1,m,C] -> np.float32 [n, m+1(+n-1)].
    Only the n x m live pairs are evaluated; padded rows/cols are exactly 0 before the softmax,
    as in the reference (AFE.py:123-127)."""
    n, m, P = xp.shape[1], xn.shape[1], max_object
    bn = lambda p, x: (x - sd[p + ".running_mean"]) / torch.sqrt(sd[p + ".running_var"] + EPS) * sd[p + ".weight"] + sd[p + ".bias"]
    u = bn("AFE.stacker2_bn", xp[0])                             # [n,C]
    v = bn("AFE.stacker2_bn", xn[0])                             # [m,C]
    z = torch.cat([u.unsqueeze(1).expand(n, m, -1), v.unsqueeze(0).expand(n, m, -1)], 2).reshape(n * m, -1)
    for idx, has_bn in ((0, True), (3, True), (6, True), (9, False), (11, False)):
        w = sd[f"AFE.final_net.{idx}.weight"]
        z = z @ w.view(w.shape[0], -1).t() + sd[f"AFE.final_net.{idx}.bias"]
        if has_bn:
            z = bn(f"AFE.final_net.{idx + 1}", z)
        z = F.relu(z)
    X = torch.zeros(P + 1, P + 1)
    X[:n, :m] = z.view(n, m)
    X[P, :] = 1.0
    X[:, P] = 1.0
    xf = F.softmax(X, dim=1)
    xt = F.softmax(X, dim=0)
    out = torch.zeros(n, m + 1)
    out[:, :m] = torch.max(xf[:n, :m], xt[:n, :m])
    out[:, m] = xf[:n, P]
    if fill_up_column and n > 1:
        out = torch.cat([out, out[:, m:m + 1].repeat(1, n - 1)], dim=1)
    return out.numpy()


def convert_detection(boxes_xyxy, h, w):
    """utils/image.py:391-412: box (x1,y1,x2,y2) in pixels -> centre in [-1,1], [1,N,1,1,2]."""
    d = np.array(boxes_xyxy, dtype=np.float64).copy()
    d[:, 2] -= d[:, 0]
    d[:, 3] -= d[:, 1]
    d[:, 0] /= w; d[:, 2] /= w; d[:, 1] /= h; d[:, 3] /= h
    c = (2 * d[:, 0:2] + d[:, 2:4]) - 1.0
    return torch.from_numpy(c.astype(float)).float().view(1, -1, 1, 1, 2)


# --------------------------------------------------------------------------- LSTM
def lstm_predict(sd, h0, c0, x, max_fut):
    """KalmanFilterLSTM.predict (kalman_filter_lstm.py:65-78) for one step, batched over T
    independent tracks: h0,c0 [T,128], x [T,F].  PyTorch gate order i,f,g,o.
    Returns hn, cn [T,128], pred [T,max_fut,4]."""
    gates = x @ sd["lstm.weight_ih_l0"].t() + sd["lstm.bias_ih_l0"] + h0 @ sd["lstm.weight_hh_l0"].t() + sd["lstm.bias_hh_l0"]
    i, f, g, o = gates.chunk(4, 1)
    cn = torch.sigmoid(f) * c0 + torch.sigmoid(i) * torch.tanh(g)
    hn = torch.sigmoid(o) * torch.tanh(cn)
    y = hn @ sd["out1.weight"].t() + sd["out1.bias"]
    y = y @ sd["out2.weight"].t() + sd["out2.bias"]
    return hn, cn, y.view(x.shape[0], max_fut, -1)
