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


# --------------------------------------------------------------------------- a9: AFE.forward / add_unmatched_dim / loss
def afe_forward(sd, fm_pre, fm_next, l_pre, l_next, max_object, false_constant=1.0):
    """AFE_module.forward (AFE.py:74-86) in eval mode: forward_selector_stacker1 on both frames, forward_stacker2
    (190-207: both sides through the SAME stacker2_bn), forward_final (209-213), add_unmatched_dim (215-233).
    Nothing is zeroed: padded centres (1.5 -> border sample) take part like real ones.
    fm_*: 13 maps [B,C,H,W]; l_*: [B,P,1,1,2] -> [B,1,P+1,P+1]."""
    P = max_object
    xp = afe_extract(sd, fm_pre, l_pre)                          # [B,P,C]
    xn = afe_extract(sd, fm_next, l_next)
    B = xp.shape[0]
    bn = lambda p, x: (x - sd[p + ".running_mean"]) / torch.sqrt(sd[p + ".running_var"] + EPS) * sd[p + ".weight"] + sd[p + ".bias"]
    out = torch.full((B, 1, P + 1, P + 1), float(false_constant))
    for b in range(B):
        u, v = bn("AFE.stacker2_bn", xp[b]), bn("AFE.stacker2_bn", xn[b])
        z = torch.cat([u.unsqueeze(1).expand(P, P, -1), v.unsqueeze(0).expand(P, P, -1)], 2).reshape(P * P, -1)
        for idx, has_bn in ((0, True), (3, True), (6, True), (9, False), (11, False)):
            w = sd[f"AFE.final_net.{idx}.weight"]
            z = z @ w.view(w.shape[0], -1).t() + sd[f"AFE.final_net.{idx}.bias"]
            if has_bn:
                z = bn(f"AFE.final_net.{idx + 1}", z)
            z = F.relu(z)
        out[b, 0, :P, :P] = z.view(P, P)
    return out


def afe_loss(inp, target, mask0, mask1, max_object):
    """AFE_module.loss (AFE.py:235-328), forward values.  inp/target [B,1,Q,Q] (Q = P+1), mask0/mask1 [B,1,Q] (0/1).
    Returns the reference's 8-tuple as Python floats + indexes_pre [B,1,P] int64."""
    P = max_object
    m0 = mask0.float().unsqueeze(3).repeat(1, 1, 1, P + 1)       # row mask      (AFE.py:238)
    m1 = mask1.float().unsqueeze(2).repeat(1, 1, P + 1, 1)       # column mask   (AFE.py:239)
    region = m0 * m1
    region_pre = region.clone(); region_pre[:, :, P, :] = 0      # 249-250
    region_next = region.clone(); region_next[:, :, :, P] = 0    # 251-252
    region_union = region_pre * region_next
    input_pre = F.softmax(region_pre * inp, dim=3)               # 255
    input_next = F.softmax(region_next * inp, dim=2)             # 256
    input_all = input_pre.clone()
    input_all[:, :, :P, :P] = ((input_pre + input_next) / 2.0)[:, :, :P, :P]
    target = target.float()
    t_pre, t_next, t_union = region_pre * target, region_next * target, region_union * target
    n_all, n_pre, n_next, n_union = target.sum(), t_pre.sum(), t_next.sum(), t_union.sum()
    loss_pre = -(t_pre * torch.log(input_pre)).sum()
    if int(n_pre.item()):
        loss_pre = loss_pre / n_pre
    loss_next = -(t_next * torch.log(input_next)).sum()
    if int(n_next.item()):
        loss_next = loss_next / n_next
    loss = -(t_pre * torch.log(input_all)).sum()
    if int(n_pre.item()) and int(n_next.item()):
        loss = loss / n_pre
    loss_sim = (t_union * torch.abs((1 - input_pre) - (1 - input_next))).sum()
    if int(n_union.item()):
        loss_sim = loss_sim / n_all
    mp = mask0[:, :, :-1] != 0
    idx_t = t_pre.max(3)[1][:, :, :-1]
    idx_pre = input_all.max(3)[1][:, :, :-1]
    npre = int(mp.sum().item())
    acc_pre = (idx_pre[mp] == idx_t[mp]).float().sum() / npre if npre else (idx_pre[mp] == idx_t[mp]).float().sum() + 1
    mn = mask1[:, :, :-1] != 0
    idx_t2 = t_next.max(2)[1][:, :, :-1]
    idx_next = input_next.max(2)[1][:, :, :-1]
    nnext = int(mn.sum().item())
    acc_next = (idx_next[mn] == idx_t2[mn]).float().sum() / nnext if nnext else (idx_next[mn] == idx_t2[mn]).float().sum() + 1
    vals = [loss_pre, loss_next, loss_sim, (loss_pre + loss_next + loss + loss_sim) / 4.0, acc_pre, acc_next, (acc_pre + acc_next) / 2.0]
    return [float(v) for v in vals], idx_pre


# --------------------------------------------------------------------------- f1: recorder decay + per-track median
# utils/tracker.py cannot be imported here (it needs `lap`, `cython_bbox`, `progress` and parses sys.argv at import: SURVEY.md
# section 8(c)) and the reference holds no test vectors for it.  PINNED nevertheless: oracle/gen_golden.py::gen_tracker_logic cuts
# FeatureRecorder, Node and STrack.get_similarity out of the reference source with `ast`, runs them unmodified on a synthetic
# 58-frame video per dataset and stores inputs + outputs in tests/golden/tracker_logic.pt; tests/test_oracle_cpu.py holds the
# two functions below to it (row selection + medians bit-exact; the decayed similarities to one float32 rounding, because
# numpy 2 promotes the reference's float32-array x np.float64-scalar product to float64 where its numpy 1.x kept float32).
def recorder_similarity(aff_by_frame, frame_index, dataset="mot", decay=1.0, decay2=0.01):
    """FeatureRecorder.update's loop (utils/tracker.py:76-90; constants 22-25): aff_by_frame {pre_index: np[n_pre, m+1]}
    (forward_stacker_features(..., fill_up_column=False)) -> {pre_index: similarity * delta}, float32 like the reference
    (numpy float32 array times a Python float)."""
    m_frame = 5 if dataset == "kitti_tracking" else (3 if dataset == "nuscenes" else 10)
    out = {}
    for pre_index, sim in aff_by_frame.items():
        d = frame_index - pre_index
        delta = pow(decay, d / 3.0) if d < m_frame else pow(decay2, d / 3.0)
        out[pre_index] = (np.asarray(sim, dtype=np.float32) * delta).astype(np.float32)
    return out


def track_similarity(nodes, frame_index, all_similarity, dataset="mot", max_track_node=50):
    """STrack.get_similarity (utils/tracker.py:219-252).  nodes: [(frame_index, det_id)] oldest first;
    all_similarity: {pre_index: np[n_pre, m+1]} of the current frame (recorder_similarity's output).
    Returns np.float32 [m+1] or None (no usable node)."""
    rows = []
    for f, i in nodes:
        if frame_index - f >= max_track_node:
            continue
        rows.append(all_similarity[f][i, :])
    if not rows:
        return None
    a = np.array(rows)
    mm = 2 if dataset == "nuscenes" else 4
    if a.shape[0] > mm:
        if a.shape[0] <= mm + 1:
            a1, a2 = a[:, : a.shape[1] - 1], np.median(a[:, -1:], axis=0)
        else:
            a1, a2 = a[a.shape[0] - mm:, : a.shape[1] - 1], np.median(a[a.shape[0] - mm:, -1:], axis=0)
        return np.concatenate((np.median(a1, axis=0), a2), axis=0)
    return np.median(a, axis=0)


# --------------------------------------------------------------------------- f2: pre-processing
def normalize_u8(img_u8_hwc, mean, std):
    """Detector.pre_process (detector.py:408-410): ((img / 255. - mean) / std).astype(float32), HWC -> [1,3,H,W]."""
    x = ((img_u8_hwc.astype(np.float64) / 255.0 - np.asarray(mean, dtype=np.float32).reshape(1, 1, 3)) /
         np.asarray(std, dtype=np.float32).reshape(1, 1, 3)).astype(np.float32)
    return torch.from_numpy(x.transpose(2, 0, 1)[None].copy())


def warp_affine_bilinear(img_u8_hwc, inv_affine, H, W):
    """Inverse-map bilinear warp with constant-0 border (the arithmetic cv2.warpAffine(INTER_LINEAR) approximates in
    1/32-pixel fixed point; detector.py:399-407).  inv_affine: 2x3, network-input (x, y, 1) -> source (x, y).
    Returns float64 [H,W,3] in [0,255]."""
    A = np.asarray(inv_affine, dtype=np.float32).reshape(2, 3)
    ys, xs = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    sx = A[0, 0] * xs + A[0, 1] * ys + A[0, 2]
    sy = A[1, 0] * xs + A[1, 1] * ys + A[1, 2]
    x0, y0 = np.floor(sx).astype(np.int64), np.floor(sy).astype(np.int64)
    lx, ly = (sx - np.floor(sx)).astype(np.float32), (sy - np.floor(sy)).astype(np.float32)
    sh, sw, _ = img_u8_hwc.shape
    img = img_u8_hwc.astype(np.float32)

    def at(yy, xx):
        ok = (yy >= 0) & (yy < sh) & (xx >= 0) & (xx < sw)
        v = img[np.clip(yy, 0, sh - 1), np.clip(xx, 0, sw - 1)]
        return v * ok[..., None]
    top = at(y0, x0) + lx[..., None] * (at(y0, x0 + 1) - at(y0, x0))
    bot = at(y0 + 1, x0) + lx[..., None] * (at(y0 + 1, x0 + 1) - at(y0 + 1, x0))
    return top + ly[..., None] * (bot - top)


# --------------------------------------------------------------------------- f3: motion fusion + linear assignment
# utils/matching.py needs `lap` (requirements.txt:14, unpinned, absent) and `cython_bbox`: it does not import.
# fuse_motion_lstm is PINNED like the f1 functions (the reference's own matching.fuse_motion + KalmanFilterLSTM.gating_distance,
# cut out with `ast` and run by gen_tracker_logic: bit-exact in float64, both observation-count branches, inf gating included).
# linear_assignment stays PARITY UNPINNED BY A RUNNING REFERENCE: lap.lapjv itself is absent; lapjv(extend_cost, cost_limit) is
# restated as the exact solution of lap's extended problem and checked against brute force (tests/test_oracle_cpu.py).
CHI2INV95_2 = 5.9915                                         # kalman_filter.py:13-15 (chi2inv95[2])


def fuse_motion_lstm(cost, pred_xy, n_obs, det_xy, lambda_=0.9):
    """matching.fuse_motion (matching.py:311-364), use_lstm=True, only_position=True, covariance = eye (tracker.py:200).
    cost [T,m] float64; pred_xy [T,2] = track.prediction_at_frame(frame_id)[:2]; n_obs [T] = len(track.observations).
    >= 300 observations: metric "maha" = squared centre distance (kalman_filter_lstm.py:95-101 with identity covariance);
    fewer: metric "gaussian", whose d = measurements[:, 3:-1] - mean[3:-1] is EMPTY for 2-vectors (lines 91-94), i.e. 0."""
    cost = np.array(cost, dtype=np.float64, copy=True)
    if cost.size == 0:
        return cost
    gating_threshold = CHI2INV95_2
    for row in range(cost.shape[0]):
        if n_obs[row] >= 300:
            d = np.asarray(det_xy, dtype=np.float64) - np.asarray(pred_xy[row], dtype=np.float64)
            gd = np.sum(d * d, axis=1)
            cost[row, gd > 5.0 * gating_threshold] = np.inf
            cost[row] = lambda_ * cost[row] + 0.05 * (1 - lambda_) * gd
        else:
            gd = np.zeros(cost.shape[1])
            cost[row, gd > 50] = np.inf
            cost[row] = lambda_ * cost[row] + 0.0005 * (1 - lambda_) * gd
    return cost


def linear_assignment(cost, thresh):
    """matching.linear_assignment (matching.py:40-55) = lap.lapjv(cost, extend_cost=True, cost_limit=thresh).  `lap`
    (requirements.txt:14, unpinned, absent here) extends the matrix to (n+m)^2 with cost_limit/2 off the real block and 0 in
    the dummy-dummy block (lap/_lapjv.pyx: `cost_c_extended[:] = cost_limit / 2.`), solves the square problem and
    reports x[i] = -1 for rows matched to a dummy.  Restated with scipy's exact solver on the same extended matrix.
    Returns (x [n] column of row i or -1, y [m] row of column j or -1)."""
    from scipy.optimize import linear_sum_assignment
    cost = np.asarray(cost, dtype=np.float64)
    n, m = cost.shape
    if cost.size == 0:
        return -np.ones(n, dtype=np.int64), -np.ones(m, dtype=np.int64)
    ext = np.full((n + m, n + m), thresh / 2.0)
    ext[n:, m:] = 0.0
    ext[:n, :m] = cost
    big = 1e9                                                # scipy rejects rows that are entirely inf; gated pairs can never win anyway
    r, c = linear_sum_assignment(np.where(np.isfinite(ext), ext, big))
    x = -np.ones(n, dtype=np.int64)
    y = -np.ones(m, dtype=np.int64)
    for i, j in zip(r, c):
        if i < n and j < m and np.isfinite(cost[i, j]):
            x[i] = j
            y[j] = i
    return x, y
