"""CPU oracle for the FGT transformer inference path (TEST INFRASTRUCTURE — never imported by the
product path; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may use it).

A functional, state_dict-driven restatement in plain PyTorch (fp32 or fp64; device-agnostic, so bench.py's
GPU-eager baseline leg can run the same code on CUDA tensors through cuBLAS/cuDNN) of
/root/reference/FGT/models/model.py and FGT/models/transformer_base/*.py. Every function cites the
reference lines it restates. Pinned against the imported reference itself by
tests/golden/make_golden.py (run in the build container, where /root/reference exists); the
resulting fixtures live in tests/golden/ and are re-checked by tests/test_oracle_golden.py.
The reference ships no tests / golden vectors of its own, so these fixtures — outputs
of the unmodified reference modules on seeded inputs — are the pin.
"""
import math

import torch
import torch.nn.functional as F

LEAK = 0.2


def _conv(x, sd, key, stride=1, pad=1, groups=1, dil=1, act=True):
    """VanillaConv.forward, FGT/models/utils/network_blocks_2d.py:37-43 (conv, then LeakyReLU(0.2))."""
    y = F.conv2d(x, sd[key + ".weight"], sd[key + ".bias"], stride=stride, padding=pad, dilation=dil,
                 groups=groups)
    return F.leaky_relu(y, LEAK) if act else y


def frame_encoder(x, sd, pre="frame_endoder.layers."):
    """Encoder.forward, FGT/models/model.py:53-66: five plain convs, then four grouped convs whose
    input is the per-group interleaved concat of x0 (output of layer 6) with the running feature."""
    bt = x.shape[0]
    y = _conv(x, sd, pre + "0", stride=2)
    y = _conv(y, sd, pre + "2")
    y = _conv(y, sd, pre + "4", stride=2)
    x0 = _conv(y, sd, pre + "6")
    y = _conv(x0, sd, pre + "8")
    h, w = x0.shape[2], x0.shape[3]
    for idx, g in ((10, 2), (12, 4), (14, 8), (16, 1)):
        a = x0.reshape(bt, g, -1, h, w)
        b = y.reshape(bt, g, -1, h, w)
        y = _conv(torch.cat([a, b], 2).reshape(bt, -1, h, w), sd, pre + str(idx), groups=g)
    return y


def flow_encoder(f, sd, pre="flow_encoder."):
    """FGT.flow_encoder, FGT/models/model.py:206-212: ReplicationPad2d(2) + 5x5, 3x3 s2, 3x3, 3x3 s2."""
    y = F.pad(f, (2, 2, 2, 2), mode="replicate")
    y = _conv(y, sd, pre + "1.featureConv", pad=0)
    y = _conv(y, sd, pre + "2.featureConv", stride=2)
    y = _conv(y, sd, pre + "3.featureConv")
    y = _conv(y, sd, pre + "4.featureConv", stride=2)
    return y


def sdpa(q, k, v):
    """Attention.forward, attention_base.py:16-22 / attention_flow.py:16-22 (dropout p=0)."""
    s = q @ k.transpose(-2, -1) / math.sqrt(q.shape[-1])
    return torch.softmax(s, dim=-1) @ v


def tmhsa(x, sd, pre, t, h, w, group=2, heads=4):
    """TMHSA.forward / .inference, attention_base.py:76-106 / 44-74: zero-pad the token grid so it
    splits into group x group zones, project Q/K/V on the padded grid, attend over all t frames of a
    zone per head, crop, output_linear."""
    bt, n, c = x.shape
    b = bt // t
    d = c // heads
    wh, ww = math.ceil(h / group), math.ceil(w / group)
    pad_b, pad_r = (wh - h % wh) % wh, (ww - w % ww) % ww
    H, W = h + pad_b, w + pad_r
    zh, zw = H // group, W // group
    g = F.pad(x.reshape(bt, h, w, c), (0, 0, 0, pad_r, 0, pad_b))
    q = F.linear(g, sd[pre + "query_embedding.weight"], sd[pre + "query_embedding.bias"])
    k = F.linear(g, sd[pre + "key_embedding.weight"], sd[pre + "key_embedding.bias"])
    v = F.linear(g, sd[pre + "value_embedding.weight"], sd[pre + "value_embedding.bias"])
    out = torch.empty(b, t, H, W, c, dtype=x.dtype, device=x.device)

    def zone(z, iy, ix):
        z = z.reshape(b, t, H, W, heads, d)[:, :, iy * zh:(iy + 1) * zh, ix * zw:(ix + 1) * zw]
        return z.permute(0, 4, 1, 2, 3, 5).reshape(b, heads, t * zh * zw, d)

    for iy in range(group):
        for ix in range(group):
            o = sdpa(zone(q, iy, ix), zone(k, iy, ix), zone(v, iy, ix))  # [b, heads, L, d]
            o = o.reshape(b, heads, t, zh, zw, d).permute(0, 2, 3, 4, 1, 5).reshape(b, t, zh, zw, c)
            out[:, :, iy * zh:(iy + 1) * zh, ix * zw:(ix + 1) * zw] = o
    out = out.reshape(bt, H, W, c)[:, :h, :w].reshape(bt, n, c)
    return F.linear(out, sd[pre + "output_linear.weight"], sd[pre + "output_linear.bias"])


def _windows(g, ws):
    """[bt, H, W, c] -> [bt, nWin, ws*ws, c] (row-major windows), attention_flow.py:132-133."""
    bt, H, W, c = g.shape
    g = g.reshape(bt, H // ws, ws, W // ws, ws, c).permute(0, 1, 3, 2, 4, 5)
    return g.reshape(bt, (H // ws) * (W // ws), ws * ws, c)


def swmhsa(x, f, sd, pre, h, w, ws=8, gd=4, heads=4):
    """SWMHSA_depthGlobalWindowConcatLN_qkFlow_reweightFlow.forward / .inference,
    attention_flow.py:115-171 / 57-113: flow re-weighting gate, window queries on [x; f'],
    keys = window tokens + depthwise-pooled global tokens (k_norm over both), values likewise from x
    (v_norm), per-window multi-head attention, un-window, crop, output_linear."""
    bt, n, c = x.shape
    cf = f.shape[2]
    d = c // heads
    pad_b, pad_r = (ws - h % ws) % ws, (ws - w % ws) % ws
    H, W = h + pad_b, w + pad_r
    xg = F.pad(x.reshape(bt, h, w, c), (0, 0, 0, pad_r, 0, pad_b))
    fg = F.pad(f.reshape(bt, h, w, cf), (0, 0, 0, pad_r, 0, pad_b))
    gate = torch.sigmoid(F.linear(torch.cat([xg, fg], -1), sd[pre + "reweightFlow.0.weight"],
                                  sd[pre + "reweightFlow.0.bias"]))
    qk = torch.cat([xg, fg * gate], -1)  # [bt, H, W, c+cf]
    ck = c + cf
    nwin = (H // ws) * (W // ws)
    q_loc = _windows(qk, ws)  # [bt, nWin, ws*ws, ck]
    kg = F.conv2d(qk.permute(0, 3, 1, 2), sd[pre + "global_extract_k.weight"], sd[pre + "global_extract_k.bias"],
                  stride=gd, groups=ck)
    kg = kg.permute(0, 2, 3, 1).reshape(bt, 1, -1, ck).expand(bt, nwin, -1, ck)
    vg = F.conv2d(xg.permute(0, 3, 1, 2), sd[pre + "global_extract_v.weight"], sd[pre + "global_extract_v.bias"],
                  stride=gd, groups=c)
    vg = vg.permute(0, 2, 3, 1).reshape(bt, 1, -1, c).expand(bt, nwin, -1, c)
    qn = F.layer_norm(q_loc, (ck,), sd[pre + "q_norm.weight"], sd[pre + "q_norm.bias"])
    kn = F.layer_norm(torch.cat([q_loc, kg], 2), (ck,), sd[pre + "k_norm.weight"], sd[pre + "k_norm.bias"])
    vn = F.layer_norm(torch.cat([_windows(xg, ws), vg], 2), (c,), sd[pre + "v_norm.weight"], sd[pre + "v_norm.bias"])
    q = F.linear(qn, sd[pre + "query_embedding.weight"], sd[pre + "query_embedding.bias"])
    k = F.linear(kn, sd[pre + "key_embedding.weight"], sd[pre + "key_embedding.bias"])
    v = F.linear(vn, sd[pre + "value_embedding.weight"], sd[pre + "value_embedding.bias"])

    def split_heads(z):
        return z.reshape(bt, nwin, -1, heads, d).transpose(2, 3)

    o = sdpa(split_heads(q), split_heads(k), split_heads(v))  # [bt, nWin, heads, ws*ws, d]
    o = o.transpose(2, 3).reshape(bt, H // ws, W // ws, ws, ws, c).permute(0, 1, 3, 2, 4, 5)
    o = o.reshape(bt, H, W, c)[:, :h, :w].reshape(bt, n, c)
    return F.linear(o, sd[pre + "output_linear.weight"], sd[pre + "output_linear.bias"])


def fusion_ffn(x, sd, pre, n_vecs, out_hw, kernel=(7, 7), stride=(3, 3), padding=(3, 3)):
    """FusionFeedForward.forward, ffn_base.py:53-77: Linear, overlap-add fold to the feature map,
    divide by the patch-coverage count, unfold back, ReLU, Linear."""
    y = F.linear(x, sd[pre + "conv1.weight"], sd[pre + "conv1.bias"])
    b, n, c = y.shape
    kk = kernel[0] * kernel[1]
    cols = y.reshape(-1, n_vecs, c).transpose(1, 2)
    img = F.fold(cols, out_hw, kernel, stride=stride, padding=padding)
    cnt = F.fold(torch.ones(cols.shape[0], kk, n_vecs, dtype=y.dtype, device=y.device), out_hw, kernel, stride=stride, padding=padding)
    y = F.unfold(img / cnt, kernel, stride=stride, padding=padding).transpose(1, 2).reshape(b, n, c)
    return F.linear(F.relu(y), sd[pre + "conv2.2.weight"], sd[pre + "conv2.2.bias"])


def temporal_block(x, sd, pre, t, h, w, out_hw):
    """TemporalTransformer.forward, model.py:124-130 (pre-norm attention, pre-norm fusion FFN)."""
    s = F.layer_norm(x, (x.shape[-1],), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"])
    x = x + tmhsa(s, sd, pre + "attention.", t, h, w)
    y = F.layer_norm(x, (x.shape[-1],), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"])
    return x + fusion_ffn(y, sd, pre + "ffn.", h * w, out_hw)


def spatial_block(x, f, sd, pre, h, w, out_hw):
    """SpatialTransformer.forward, model.py:144-149 (attention on raw x, pre-norm fusion FFN)."""
    x = x + swmhsa(x, f, sd, pre + "attention.", h, w)
    y = F.layer_norm(x, (x.shape[-1],), sd[pre + "norm.weight"], sd[pre + "norm.bias"])
    return x + fusion_ffn(y, sd, pre + "ffn.", h * w, out_hw)


def add_pos_emb(x, sd, h, w, pre="add_pos_emb.proj"):
    """AddPosEmb.forward, model.py:75-88: depthwise 3x3 conv on the token grid + identity."""
    bt, n, c = x.shape
    g = x.transpose(1, 2).reshape(bt, c, h, w)
    g = F.conv2d(g, sd[pre + ".weight"], sd[pre + ".bias"], padding=1, groups=c) + g
    return g.flatten(2).transpose(1, 2)


def vec2patch(x, sd, out_hw, kernel=(7, 7), stride=(3, 3), padding=(3, 3), pre="vec2patch.embedding"):
    """Vec2Patch.forward, model.py:102-110: Linear to 49*C then overlap-add fold."""
    y = F.linear(x, sd[pre + ".weight"], sd[pre + ".bias"]).transpose(1, 2)
    return F.fold(y, out_hw, kernel, stride=stride, padding=padding)


def _up2(x):
    """VanillaDeconv.forward, network_blocks_2d.py:58-60: nearest x2 then conv."""
    return F.interpolate(x, scale_factor=2)


def decoder(x, sd, pre="decoder."):
    """Decoder.forward, model.py:188-193."""
    y = _conv(_up2(x), sd, pre + "layer1.conv.featureConv")
    y = _conv(y, sd, pre + "layer2.featureConv")
    y = _conv(_up2(y), sd, pre + "layer3.conv.featureConv")
    return _conv(y, sd, pre + "final.featureConv", act=False)


def fgt_forward(sd, frames, flows, masks, num_blocks=8, kernel=(7, 7), stride=(3, 3), padding=(3, 3),
                return_intermediates=False):
    """FGT.forward, model.py:249-283. `sd` is Model.state_dict() with the leading 'net.' stripped.
    frames [b,t,3,H,W], flows [b,t,2,H,W], masks [b,t,1,H,W] -> [b*t,3,H,W]."""
    b, t, _, H, W = frames.shape
    x = torch.cat([frames, masks], 2).reshape(b * t, 4, H, W)
    fl = flows.reshape(b * t, 2, H, W)
    enc = frame_encoder(x, sd)
    fenc = flow_encoder(fl, sd)
    out_hw = (H // 4, W // 4)
    tok = F.conv2d(enc, sd["patch2vec.weight"], sd["patch2vec.bias"], stride=stride, padding=padding)
    ftok = F.conv2d(fenc, sd["f_patch2vec.weight"], sd["f_patch2vec.bias"], stride=stride, padding=padding)
    h, w = tok.shape[2], tok.shape[3]
    tok = tok.flatten(2).transpose(1, 2)
    ftok = ftok.flatten(2).transpose(1, 2)
    inter = {"enc": enc, "tok0": tok, "ftok": ftok}
    tok = temporal_block(tok, sd, "first_t_transformer.", t, h, w, out_hw)
    inter["t0"] = tok
    tok = add_pos_emb(tok, sd, h, w)
    tok = spatial_block(tok, ftok, sd, "first_s_transformer.", h, w, out_hw)
    inter["s0"] = tok
    for i in range(num_blocks // 2 - 1):
        tok = temporal_block(tok, sd, f"transformer.{i}.t_transformer.", t, h, w, out_hw)
        tok = spatial_block(tok, ftok, sd, f"transformer.{i}.s_transformer.", h, w, out_hw)
    inter["tok_final"] = tok
    feat = enc + vec2patch(tok, sd, out_hw, kernel, stride, padding)
    out = torch.tanh(decoder(feat, sd))
    if return_intermediates:
        return out, inter
    return out


def strip_net(state_dict):
    """Model.state_dict() keys are 'net.*' (model.py:15); the oracle uses the inner names."""
    return {k[4:] if k.startswith("net.") else k: v for k, v in state_dict.items()}
