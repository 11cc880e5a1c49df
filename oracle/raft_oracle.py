"""CPU oracle for RAFT optical flow (TEST INFRASTRUCTURE — see oracle/fgt_oracle.py header).

Functional PyTorch-CPU restatement of /root/reference/RAFT/raft.py:87-145 (RAFT.forward, basic
model, eval mode), RAFT/extractor.py:6-56,118-192 (BasicEncoder / ResidualBlock), RAFT/corr.py:12-60
(CorrBlock), RAFT/update.py (BasicMotionEncoder :79-97, SepConvGRU :33-60, FlowHead :6-14,
BasicUpdateBlock :114-136) and RAFT/utils/utils.py:57-76. Pinned by tests/golden/raft_*.npz (outputs
of the unmodified reference) and, in the build container, against the real raft-things.pth weights
(tests/golden/make_golden.py).
"""
import torch
import torch.nn.functional as F


def _norm(x, sd, key, kind):
    if kind == "instance":  # nn.InstanceNorm2d: no affine, per-image statistics (extractor.py:29-33)
        return F.instance_norm(x, eps=1e-5)
    return F.batch_norm(x, sd[key + ".running_mean"], sd[key + ".running_var"], sd[key + ".weight"],
                        sd[key + ".bias"], training=False, eps=1e-5)  # eval-mode BatchNorm2d


def _conv(x, sd, key, stride=1, pad=0):
    return F.conv2d(x, sd[key + ".weight"], sd[key + ".bias"], stride=stride, padding=pad)


def res_block(x, sd, key, kind, stride):
    """ResidualBlock.forward, extractor.py:47-56."""
    y = F.relu(_norm(_conv(x, sd, key + ".conv1", stride, 1), sd, key + ".norm1", kind))
    y = F.relu(_norm(_conv(y, sd, key + ".conv2", 1, 1), sd, key + ".norm2", kind))
    if stride != 1:
        x = _norm(_conv(x, sd, key + ".downsample.0", stride, 0), sd, key + ".norm3", kind)
    return F.relu(x + y)


def encoder(x, sd, key, kind):
    """BasicEncoder.forward, extractor.py:168-192."""
    x = F.relu(_norm(_conv(x, sd, key + ".conv1", 2, 3), sd, key + ".norm1", kind))
    for name, stride in (("layer1", 1), ("layer2", 2), ("layer3", 2)):
        x = res_block(x, sd, f"{key}.{name}.0", kind, stride)
        x = res_block(x, sd, f"{key}.{name}.1", kind, 1)
    return _conv(x, sd, key + ".conv2")


def corr_pyramid(f1, f2, levels=4):
    """CorrBlock.__init__ / corr, corr.py:13-27,52-60."""
    b, d, h, w = f1.shape
    corr = torch.matmul(f1.reshape(b, d, h * w).transpose(1, 2), f2.reshape(b, d, h * w))
    corr = (corr / torch.sqrt(torch.tensor(d).float())).reshape(b * h * w, 1, h, w)
    pyr = [corr]
    for _ in range(levels - 1):
        corr = F.avg_pool2d(corr, 2, stride=2)
        pyr.append(corr)
    return pyr


def sample(img, coords):
    """bilinear_sampler, utils/utils.py:57-71 (pixel coordinates, align_corners=True, zero padding)."""
    H, W = img.shape[-2:]
    x, y = coords.split([1, 1], dim=-1)
    grid = torch.cat([2 * x / (W - 1) - 1, 2 * y / (H - 1) - 1], dim=-1)
    return F.grid_sample(img, grid, align_corners=True)


def corr_lookup(pyr, coords, r=4):
    """CorrBlock.__call__, corr.py:29-50."""
    coords = coords.permute(0, 2, 3, 1)
    b, h, w, _ = coords.shape
    out = []
    d = torch.linspace(-r, r, 2 * r + 1)
    delta = torch.stack(torch.meshgrid(d, d, indexing="ij"), dim=-1).view(1, 2 * r + 1, 2 * r + 1, 2)
    for i, corr in enumerate(pyr):
        c = coords.reshape(b * h * w, 1, 1, 2) / 2 ** i + delta
        out.append(sample(corr, c).view(b, h, w, -1))
    return torch.cat(out, dim=-1).permute(0, 3, 1, 2).contiguous().float()


def motion_encoder(flow, corr, sd, key="update_block.encoder"):
    """BasicMotionEncoder.forward, update.py:89-97."""
    cor = F.relu(_conv(corr, sd, key + ".convc1"))
    cor = F.relu(_conv(cor, sd, key + ".convc2", 1, 1))
    flo = F.relu(_conv(flow, sd, key + ".convf1", 1, 3))
    flo = F.relu(_conv(flo, sd, key + ".convf2", 1, 1))
    out = F.relu(_conv(torch.cat([cor, flo], 1), sd, key + ".conv", 1, 1))
    return torch.cat([out, flow], 1)


def sep_conv_gru(h, x, sd, key="update_block.gru"):
    """SepConvGRU.forward, update.py:45-60."""
    for sfx, pad in (("1", (0, 2)), ("2", (2, 0))):
        hx = torch.cat([h, x], 1)
        z = torch.sigmoid(F.conv2d(hx, sd[f"{key}.convz{sfx}.weight"], sd[f"{key}.convz{sfx}.bias"], padding=pad))
        r = torch.sigmoid(F.conv2d(hx, sd[f"{key}.convr{sfx}.weight"], sd[f"{key}.convr{sfx}.bias"], padding=pad))
        q = torch.tanh(F.conv2d(torch.cat([r * h, x], 1), sd[f"{key}.convq{sfx}.weight"],
                                sd[f"{key}.convq{sfx}.bias"], padding=pad))
        h = (1 - z) * h + z * q
    return h


def upsample_flow(flow, mask):
    """RAFT.upsample_flow, raft.py:73-84."""
    n, _, h, w = flow.shape
    mask = torch.softmax(mask.view(n, 1, 9, 8, 8, h, w), dim=2)
    up = F.unfold(8 * flow, [3, 3], padding=1).view(n, 2, 9, 1, 1, h, w)
    up = torch.sum(mask * up, dim=2).permute(0, 1, 4, 2, 5, 3)
    return up.reshape(n, 2, 8 * h, 8 * w)


def raft_forward(sd, image1, image2, iters=20, hdim=128, cdim=128, all_iters=False):
    """RAFT.forward(test_mode=True), raft.py:87-145. sd: state_dict without the 'module.' prefix.
    Returns (flow_low [n,2,H/8,W/8], flow_up [n,2,H,W]); with all_iters also the per-iteration
    low-resolution flows."""
    im1 = (2 * (image1 / 255.0) - 1.0).contiguous()
    im2 = (2 * (image2 / 255.0) - 1.0).contiguous()
    fm = encoder(torch.cat([im1, im2], 0), sd, "fnet", "instance")
    n = im1.shape[0]
    f1, f2 = fm[:n].float(), fm[n:].float()
    pyr = corr_pyramid(f1, f2)
    c = encoder(im1, sd, "cnet", "batch")
    net, inp = torch.tanh(c[:, :hdim]), torch.relu(c[:, hdim:hdim + cdim])
    _, _, H, W = im1.shape
    ys, xs = torch.meshgrid(torch.arange(H // 8), torch.arange(W // 8), indexing="ij")
    coords0 = torch.stack([xs, ys], 0).float()[None].repeat(n, 1, 1, 1)
    coords1 = coords0.clone()
    lows = []
    flow_up = None
    for _ in range(iters):
        corr = corr_lookup(pyr, coords1)
        flow = coords1 - coords0
        mf = motion_encoder(flow, corr, sd)
        net = sep_conv_gru(net, torch.cat([inp, mf], 1), sd)
        delta = _conv(F.relu(_conv(net, sd, "update_block.flow_head.conv1", 1, 1)), sd,
                      "update_block.flow_head.conv2", 1, 1)
        mask = 0.25 * _conv(F.relu(_conv(net, sd, "update_block.mask.0", 1, 1)), sd, "update_block.mask.2")
        coords1 = coords1 + delta
        flow_up = upsample_flow(coords1 - coords0, mask)
        lows.append(coords1 - coords0)
    if all_iters:
        return coords1 - coords0, flow_up, lows
    return coords1 - coords0, flow_up
