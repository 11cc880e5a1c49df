"""B200-native RAFT (basic model) behind the reference's nn.Module interface.

Drop-in for /root/reference/RAFT/raft.py: `RAFT(args).forward(image1, image2, iters=12,
flow_init=None, upsample=True, test_mode=False)` with the reference's state_dict keys (179 tensors of
raft-things.pth; survives the driver's DataParallel wrap/unwrap, tool/video_inpainting.py:186-197,
call site :263). Parameter holders only; arithmetic runs in libfgt_sm100a.so:

  * feature / context encoders: implicit-GEMM convs; InstanceNorm = fgt_chan_stats + fgt_instnorm_act,
    eval-mode BatchNorm folded into the conv weights at pack time;
  * all-pairs correlation: one tcgen05 GEMM fmap1 x fmap2^T (alpha = 1/sqrt(256)) + fgt_avgpool2 pyramid;
  * per iteration: fgt_corr_lookup -> motion encoder convs -> SepConvGRU (z, r*h, and the
    (1-z)h+zq update fused in GEMM epilogues) -> flow head -> fgt_raft_flow_update;
  * mask head + fgt_convex_upsample only where the result is consumed (last iteration in test_mode;
    the reference computes and discards the other 19, RAFT/raft.py:134-143).
"""
import torch
import torch.nn as nn

from . import lib, ops


class _ResBlock(nn.Module):
    def __init__(self, cin, planes, norm, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, planes, 3, padding=1, stride=stride)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1)
        mk = (lambda: nn.BatchNorm2d(planes)) if norm == "batch" else (lambda: nn.InstanceNorm2d(planes))
        self.norm1, self.norm2 = mk(), mk()
        self.stride = stride
        if stride != 1:
            self.norm3 = mk()
            self.downsample = nn.Sequential(nn.Conv2d(cin, planes, 1, stride=stride), self.norm3)


class _Encoder(nn.Module):
    def __init__(self, out_dim, norm):
        super().__init__()
        self.norm_fn = norm
        self.norm1 = nn.BatchNorm2d(64) if norm == "batch" else nn.InstanceNorm2d(64)
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3)
        self.layer1 = nn.Sequential(_ResBlock(64, 64, norm, 1), _ResBlock(64, 64, norm, 1))
        self.layer2 = nn.Sequential(_ResBlock(64, 96, norm, 2), _ResBlock(96, 96, norm, 1))
        self.layer3 = nn.Sequential(_ResBlock(96, 128, norm, 2), _ResBlock(128, 128, norm, 1))
        self.conv2 = nn.Conv2d(128, out_dim, 1)


class _MotionEncoder(nn.Module):
    def __init__(self):
        super().__init__()
        self.convc1 = nn.Conv2d(324, 256, 1)
        self.convc2 = nn.Conv2d(256, 192, 3, padding=1)
        self.convf1 = nn.Conv2d(2, 128, 7, padding=3)
        self.convf2 = nn.Conv2d(128, 64, 3, padding=1)
        self.conv = nn.Conv2d(256, 126, 3, padding=1)


class _GRU(nn.Module):
    def __init__(self):
        super().__init__()
        for g in "zrq":
            setattr(self, f"conv{g}1", nn.Conv2d(384, 128, (1, 5), padding=(0, 2)))
        for g in "zrq":
            setattr(self, f"conv{g}2", nn.Conv2d(384, 128, (5, 1), padding=(2, 0)))


class _FlowHead(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(128, 256, 3, padding=1)
        self.conv2 = nn.Conv2d(256, 2, 3, padding=1)


class _UpdateBlock(nn.Module):
    def __init__(self):
        super().__init__()
        self.encoder = _MotionEncoder()
        self.gru = _GRU()
        self.flow_head = _FlowHead()
        self.mask = nn.Sequential(nn.Conv2d(128, 256, 3, padding=1), nn.ReLU(), nn.Conv2d(256, 576, 1))


class RAFT(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.args = args
        if getattr(args, "small", False):
            raise ValueError("fgt_b200 RAFT implements the basic model (the driver never sets --small)")
        if getattr(args, "alternate_corr", False):
            raise ValueError("alternate_corr is unreachable in the reference (RAFT/raft.py:106) and unsupported")
        self.hidden_dim = self.context_dim = 128
        args.corr_levels, args.corr_radius = 4, 4
        self.fnet = _Encoder(256, "instance")
        self.cnet = _Encoder(256, "batch")
        self.update_block = _UpdateBlock()
        for net in (self.fnet, self.cnet):  # extractor.py:146-153
            for m in net.modules():
                if isinstance(m, nn.Conv2d):
                    nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
        self._packed = None
        self._bufs = {}

    def _drop_graphs(self):
        """Captured CUDA graphs hold the addresses of the packed weights / workspaces they were recorded with."""
        if getattr(self, "_graphs", None) is not None:
            self._graphs.clear()

    def _apply(self, fn, *a, **k):
        self._packed = None
        self._bufs = {}
        self._drop_graphs()
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._packed = None
        self._drop_graphs()
        return super().load_state_dict(*a, **k)

    # ------------------------------------------------------------------ weight packing
    def _pack(self, dev):
        sd = {k: v.detach().float() for k, v in self.state_dict().items() if v.dtype.is_floating_point}
        P = {}

        def wb(key, bn=None):
            w, b = sd[key + ".weight"], sd[key + ".bias"]
            if bn is not None:  # eval-mode BatchNorm folded into the conv (raft.py:55, model.eval())
                g, be = sd[bn + ".weight"].double(), sd[bn + ".bias"].double()
                mu, var = sd[bn + ".running_mean"].double(), sd[bn + ".running_var"].double()
                s = g / torch.sqrt(var + 1e-5)
                w = (w.double() * s[:, None, None, None]).float()
                b = ((b.double() - mu) * s + be).float()
            return w, b

        def put(name, key, bn=None, segs=None, im2col=None):
            w, b = wb(key, bn)
            P[name] = ops.packed(name, w, b, dev, segs, im2col)

        for net, isbn in (("fnet", False), ("cnet", True)):
            put(f"{net}.conv1", f"{net}.conv1", f"{net}.norm1" if isbn else None, im2col=192)
            for li in (1, 2, 3):
                for bi in (0, 1):
                    k = f"{net}.layer{li}.{bi}"
                    put(k + ".conv1", k + ".conv1", k + ".norm1" if isbn else None)
                    put(k + ".conv2", k + ".conv2", k + ".norm2" if isbn else None)
                    if li > 1 and bi == 0:
                        put(k + ".down", k + ".downsample.0", k + ".norm3" if isbn else None)
            put(f"{net}.conv2", f"{net}.conv2")
        u = "update_block."
        w, b = wb(u + "encoder.convc1")
        w = torch.nn.functional.pad(w, (0, 0, 0, 0, 0, 384 - 324))  # lookup rows are 384 wide (324 + zero pad)
        P["convc1"] = ops.packed("convc1", w, b, dev)
        put("convc2", u + "encoder.convc2")
        put("convf1", u + "encoder.convf1", im2col=128)
        put("convf2", u + "encoder.convf2")
        put("mconv", u + "encoder.conv")
        for s in "12":
            put(f"gru.q{s}", u + f"gru.convq{s}", segs=[128, 256])
            (wz, bz), (wr, br) = wb(u + f"gru.convz{s}", None), wb(u + f"gru.convr{s}", None)
            P[f"gru.zr{s}"] = ops.packed(f"gru.zr{s}", torch.cat([wz, wr], 0), torch.cat([bz, br], 0), dev, [128, 256])
        put("fh1", u + "flow_head.conv1")
        put("fh2", u + "flow_head.conv2")
        w2 = sd[u + "flow_head.conv2.weight"]
        P["fh2t"] = ops.packed("fh2t", lib.pack_taps_as_n(w2), torch.zeros(32), dev)
        put("mask0", u + "mask.0")
        put("mask2", u + "mask.2")
        P["mask2"]["b"] = P["mask2"]["b"] * 0.25  # mask = 0.25 * conv(x) (update.py:135): alpha scales W x
        self._packed = P
        return P

    def _buf(self, key, name, shape, dev, split=True, zero=False, dtype=None):
        d = self._bufs.setdefault(key, {})
        if name not in d:
            if split:
                shape, dtype = (2,) + tuple(shape), torch.bfloat16
            elif dtype is None:
                dtype = torch.float32
            d[name] = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=dev)
        return d[name]

    # ------------------------------------------------------------------ encoders
    def _encoder(self, img, net, P, dev, n):
        """BasicEncoder.forward (extractor.py:168-192) on `n` raw [0,255] images [n,3,H,W] -> NHWC features."""
        _, _, H, W = img.shape
        inst = net == "fnet"
        key = (net, n, H, W)
        B = lambda nm, s, **kw: self._buf(key, nm, s, dev, **kw)  # noqa: E731
        RELU = lib.ACT_RELU
        H2, W2 = H // 2, W // 2
        col = B("col", (n, H2, W2, 192))
        # 2*(x/255)-1 applied to in-bounds taps (raft.py:90-91); zero padding stays zero
        lib.im2col_nchw(img, None, col, k=7, stride=2, pad=3, replicate=False, OH=H2, OW=W2, scale=2.0 / 255.0,
                        shift=-1.0, tag=net + ".conv1")
        stats = B("stats", (n * 128 * 2,), split=False, dtype=torch.float64)
        raw = B("raw", (n * H2 * W2 * 64,), split=False)  # largest pre-norm conv output, reused

        def conv_norm_act(xs, name, geo_n, hh, ww, c_out, stride, out_split, out_f32=None, res=None, relu=True, k=3):
            """conv (+folded BN) + norm + relu [+ residual relu]; returns nothing, writes outputs."""
            oh, ow = (hh + 2 * (k // 2) - k) // stride + 1, (ww + 2 * (k // 2) - k) // stride + 1
            if inst:
                r = raw[: geo_n * oh * ow * c_out].view(geo_n, oh, ow, c_out)
                ops.conv(xs, P[name], kx=k, ky=k, stride=stride, pad_x=k // 2, pad_y=k // 2, act=lib.ACT_NONE, out_f32=r)
                lib.chan_stats(r, geo_n, oh * ow, c_out, stats)
                lib.instnorm_act(r, stats, geo_n, oh * ow, c_out, relu=relu, res=res, out=out_f32, out_split=out_split)
            else:
                if res is not None:
                    ops.conv(xs, P[name], kx=k, ky=k, stride=stride, pad_x=k // 2, pad_y=k // 2, act=RELU, aux=res,
                             aux_mode=lib.AUX_ADD_RELU, out_split=out_split, out_f32=out_f32)
                else:
                    ops.conv(xs, P[name], kx=k, ky=k, stride=stride, pad_x=k // 2, pad_y=k // 2,
                             act=RELU if relu else lib.ACT_NONE, out_split=out_split, out_f32=out_f32)

        x = B("x0", (n, H2, W2, 64))
        xf = B("x0f", (n, H2, W2, 64), split=False)
        if inst:
            r = raw[: n * H2 * W2 * 64].view(n, H2, W2, 64)
            ops.linear([(col.view(2, n * H2 * W2, 192), 192)], P[net + ".conv1"], n * H2 * W2, out_f32=r)
            lib.chan_stats(r, n, H2 * W2, 64, stats)
            lib.instnorm_act(r, stats, n, H2 * W2, 64, relu=True, out=xf, out_split=x)
        else:
            ops.linear([(col.view(2, n * H2 * W2, 192), 192)], P[net + ".conv1"], n * H2 * W2, act=RELU, out_f32=xf,
                       out_split=x)
        hh, ww, cin = H2, W2, 64
        for li, (dim, stride) in enumerate(((64, 1), (96, 2), (128, 2)), start=1):
            for bi in (0, 1):
                st = stride if bi == 0 else 1
                k = f"{net}.layer{li}.{bi}"
                oh, ow = hh // st, ww // st
                y1 = B(f"{k}.y1", (n, oh, ow, dim))
                conv_norm_act([(x, cin)], k + ".conv1", n, hh, ww, dim, st, y1)
                if st != 1:  # x = norm3(conv1x1_s2(x)), no ReLU (extractor.py:52-53)
                    ds = B(f"{k}.ds", (n, oh, ow, dim), split=False)
                    conv_norm_act([(x, cin)], k + ".down", n, hh, ww, dim, st, None, out_f32=ds, relu=False, k=1)
                    res = ds
                else:
                    res = xf
                xo = B(f"{k}.out", (n, oh, ow, dim))
                xof = B(f"{k}.outf", (n, oh, ow, dim), split=False)
                conv_norm_act([(y1, dim)], k + ".conv2", n, oh, ow, dim, 1, xo, out_f32=xof, res=res)
                x, xf, hh, ww, cin = xo, xof, oh, ow, dim
        return x, hh, ww  # NHWC split [2, n, H/8, W/8, 128]

    # ------------------------------------------------------------------ forward
    def _forward_batch(self, im1, im2, iters, flow_init, test_mode, P, dev):
        """RAFT.forward (raft.py:86-148) on n image pairs at once: every kernel of the update loop covers
        the n*h*w positions of all pairs, which is what fills the 148 SMs (one 1/8-resolution pair is 51 tiles)."""
        n, _, H, W = im1.shape
        if H % 8 or W % 8:
            raise ValueError(f"RAFT input {H}x{W} must be divisible by 8 (the reference crashes otherwise, "
                             "RAFT/raft.py:64-71 vs extractor.py)")
        h, w = H // 8, W // 8
        npx = h * w
        tot = n * npx
        key = ("iter", n, H, W)
        B = lambda nm, s, **kw: self._buf(key, nm, s, dev, **kw)  # noqa: E731
        RELU, SIG, TANH, NONE = lib.ACT_RELU, lib.ACT_SIGMOID, lib.ACT_TANH, lib.ACT_NONE
        # ---- feature net on all 2n images, 1x1 output conv -> fmaps [2, 2n*npx, 256] (split)
        f128, _, _ = self._encoder(torch.cat([im1, im2], 0), "fnet", P, dev, 2 * n)
        fmap = B("fmap", (2 * tot, 256))
        ops.linear([(f128.view(2, 2 * tot, 128), 128)], P["fnet.conv2"], 2 * tot, out_split=fmap)
        # ---- all-pairs correlation (corr.py:52-60): corr[i,j] = <f1_i, f2_j> / 16 per pair, then the pyramid
        pyr = [B("corr0", (tot, h, w), split=False)]
        for i in range(n):
            f1 = lib.ASeg(fmap[:, i * npx:(i + 1) * npx], 256, npx)
            f2 = fmap[:, (n + i) * npx:(n + i + 1) * npx]  # rows of image 2 as the "weight" operand [N=npx, K=256]
            lib.gemm_tc([f1], f2, npx, out_w=npx, bn=128, alpha=1.0 / 16.0, out_f32=pyr[0][i * npx:(i + 1) * npx],
                        tag="corr")
        hh, ww = h, w
        for i in range(1, 4):
            nxt = B(f"corr{i}", (tot, hh // 2, ww // 2), split=False)
            lib.avgpool2(pyr[-1], tot, hh, ww, nxt)
            pyr.append(nxt)
            hh, ww = hh // 2, ww // 2
        # ---- context net: net = tanh(c[:128]), inp = relu(c[128:]) (raft.py:112-115)
        c128, _, _ = self._encoder(im1, "cnet", P, dev, n)
        hsp = B("h", (tot, 128))
        hf = B("hf", (tot, 128), split=False)
        xbuf = B("x", (tot, 256))  # GRU input x = [inp(128) | motion features(126) | flow(2)]
        cw = P["cnet.conv2"]
        if "cnet.net" not in P:
            P["cnet.net"] = dict(cw, w=cw["w"][:, :128].contiguous(), b=cw["b"][:128].contiguous(), N=128, name="cnet.net")
            P["cnet.inp"] = dict(cw, w=cw["w"][:, 128:].contiguous(), b=cw["b"][128:].contiguous(), N=128, name="cnet.inp")
        ops.linear([(c128.view(2, tot, 128), 128)], P["cnet.net"], tot, act=TANH, out_split=hsp, out_f32=hf)
        ops.linear([(c128.view(2, tot, 128), 128)], P["cnet.inp"], tot, act=RELU, out_split=xbuf, os_x=256)
        # ---- iterations
        coords = B("coords", (tot, 2), split=False)
        ys, xs = torch.meshgrid(torch.arange(h, device=dev), torch.arange(w, device=dev), indexing="ij")
        coords.view(n, npx, 2).copy_(torch.stack([xs, ys], -1).reshape(1, npx, 2).float().expand(n, npx, 2))
        if flow_init is not None:
            coords.add_(flow_init.permute(0, 2, 3, 1).reshape(tot, 2).float())
        flow_nchw = B("flow", (n, 2, h, w), split=False)
        lib.raft_flow_update(coords, None, h, w, flow_nchw, xbuf, 254, n=n)
        look = B("look", (tot, 384), zero=True)
        cor1 = B("cor1", (n, h, w, 256))
        corflo = B("corflo", (n, h, w, 256))  # [cor(192) | flo(64)]
        fcol = B("fcol", (n, h, w, 128))
        flo1 = B("flo1", (n, h, w, 128))
        z = B("z", (tot, 128), split=False)
        rh = B("rh", (n, h, w, 128))
        fh = B("fh", (n, h, w, 256))
        delta = B("delta", (tot, 2), split=False)
        fh2y = B("fh2y", (32, tot), split=False)  # column-planar partial products of the flow head
        m0 = B("m0", (n, h, w, 256))
        mask = B("mask", (tot, 576), split=False)
        h4 = hsp.view(2, n, h, w, 128)
        x4 = xbuf.view(2, n, h, w, 256)
        ups = []
        for it in range(iters):
            lib.corr_lookup(pyr, coords, tot, 4, look)
            # motion encoder (update.py:89-97)
            ops.linear([(look, 384)], P["convc1"], tot, act=RELU, out_split=cor1.view(2, tot, 256))
            ops.conv([(cor1, 256)], P["convc2"], kx=3, ky=3, pad_x=1, pad_y=1, act=RELU, out_split=corflo,
                     out_c_total=256, out_c_offset=0)
            lib.im2col_nchw(flow_nchw, None, fcol, k=7, stride=1, pad=3, replicate=False, OH=h, OW=w)
            ops.linear([(fcol.view(2, tot, 128), 128)], P["convf1"], tot, act=RELU, out_split=flo1.view(2, tot, 128))
            ops.conv([(flo1, 128)], P["convf2"], kx=3, ky=3, pad_x=1, pad_y=1, act=RELU, out_split=corflo,
                     out_c_total=256, out_c_offset=192)
            ops.conv([(corflo, 256)], P["mconv"], kx=3, ky=3, pad_x=1, pad_y=1, act=RELU, out_split=x4,
                     out_c_total=256, out_c_offset=128)
            # SepConvGRU (update.py:45-60): horizontal (1x5) then vertical (5x1)
            for s, (kx, ky) in (("1", (5, 1)), ("2", (1, 5))):
                kw = dict(kx=kx, ky=ky, pad_x=kx // 2, pad_y=ky // 2, seg_counts=[128, 256])
                # z = sigmoid(convz(hx)) and r*h = sigmoid(convr(hx)) * h as ONE N=256 GEMM over the shared input
                ops.conv([(h4, 128), (x4, 256)], P[f"gru.zr{s}"], act=SIG, aux=hf, aux_mode=lib.AUX_GRU_ZR,
                         out_f32=z.view(n, h, w, 128), out_split=rh, out_c_total=128, **kw)
                ops.conv([(rh, 128), (x4, 256)], P[f"gru.q{s}"], act=TANH, aux=hf, aux2=z, aux_mode=lib.AUX_GRU,
                         out_f32=hf.view(n, h, w, 128), out_split=h4, **kw)
            # flow head (update.py:13-14)
            ops.conv([(h4, 128)], P["fh1"], kx=3, ky=3, pad_x=1, pad_y=1, act=RELU, out_split=fh)
            # 256->2 conv as "taps as N" (1x1 GEMM with N = 18 -> 32, then shift-and-add): A is read once, not 9x
            ops.linear([(fh.view(2, tot, 256), 256)], P["fh2t"], tot, out_f32=fh2y, os_x=1, os_c=tot)
            lib.tapsum(fh2y, n, h, w, 2, 3, P["fh2"]["b"], NONE, delta, nchw=False, tag="fh2")
            lib.raft_flow_update(coords, delta, h, w, flow_nchw, xbuf, 254, n=n)
            if not test_mode or it == iters - 1:
                # mask head scaled by 0.25 (update.py:122-125,135) + convex upsampling (raft.py:73-84)
                ops.conv([(h4, 128)], P["mask0"], kx=3, ky=3, pad_x=1, pad_y=1, act=RELU, out_split=m0)
                ops.linear([(m0.view(2, tot, 256), 256)], P["mask2"], tot, alpha=0.25, out_f32=mask)
                up = torch.empty(n, 2, 8 * h, 8 * w, device=dev)
                lib.convex_upsample(mask, flow_nchw, h, w, up, n=n)
                ups.append(up)
        return flow_nchw.clone(), ups

    def enable_cuda_graph(self, on=True):
        """Replay whole forwards as CUDA graphs, one per (geometry, iters, test_mode) (fgt_b200/graphs.py)."""
        self._graphs = {} if on else None

    def forward(self, image1, image2, iters=12, flow_init=None, upsample=True, test_mode=False):
        if not image1.is_cuda:
            raise RuntimeError("fgt_b200 RAFT runs on a CUDA (sm_100a) device only; there is no CPU fallback")
        if getattr(self, "_graphs", None) is not None and flow_init is None:
            from .graphs import GraphedCall
            key = (iters, bool(test_mode))
            if key not in self._graphs:
                self._graphs[key] = GraphedCall(
                    lambda a, b, _it=iters, _tm=test_mode: self._forward_impl(a, b, _it, None, _tm))
            out = self._graphs[key](image1.float().contiguous(), image2.float().contiguous())
            return list(out) if not test_mode else out
        return self._forward_impl(image1, image2, iters, flow_init, test_mode)

    def _forward_impl(self, image1, image2, iters, flow_init, test_mode):
        dev = image1.device
        P = self._packed if self._packed is not None else self._pack(dev)
        image1, image2 = image1.float().contiguous(), image2.float().contiguous()
        lo, ups = self._forward_batch(image1, image2, iters, flow_init, test_mode, P, dev)
        if test_mode:
            return lo, ups[-1]
        return ups
