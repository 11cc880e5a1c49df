"""GPU parity of the individual sm_100a kernels, called through the C-ABI (fgt_b200.lib), against
fp64 PyTorch restatements of the same op on the same seeded inputs."""

import pytest
import torch
import torch.nn.functional as F

from tests.util import assert_close

pytestmark = pytest.mark.gpu

KTOL = 1e-4  # split-bf16 3-term products: ~1e-5 observed, fp32-grade


def _lib():
    from fgt_b200 import lib
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    lib.load()
    return lib


@pytest.mark.parametrize("M,N,K,bn,act,aux,split", [
    (128, 128, 64, 128, 0, 0, False), (256, 256, 512, 128, 1, 0, True), (1000, 520, 1960, 128, 2, 0, True),
    (7200, 1536, 512, 256, 0, 1, False), (300, 48, 200, 48, 3, 2, False), (40000, 512, 512, 128, 4, 0, False),
])
def test_linear(M, N, K, bn, act, aux, split):
    from tools import diag_gemm as D
    assert D.linear_case(M, N, K, bn=bn, act=act, aux_mode=aux, split_out=split)


@pytest.mark.parametrize("kw", [
    dict(n=1, h=16, w=32, cin=64, cout=64, k=3), dict(n=2, h=60, w=108, cin=128, cout=256, k=3, bn=128),
    dict(n=2, h=61, w=107, cin=64, cout=128, k=3, stride=2, bn=128),
    dict(n=1, h=60, w=108, cin=128, cout=512, k=7, stride=3, pad=3, bn=128, act=0),
    dict(n=1, h=30, w=54, cin=192, cout=192, k=3, dil=4, bn=64),
    dict(n=2, h=24, w=40, cin=128, cout=256, k=3, groups=2, bn=128),
    dict(n=1, h=24, w=40, cin=8, cout=64, k=3, bn=64), dict(n=1, h=20, w=36, cin=64, cout=3, k=3, bn=16, act=0),
])
def test_conv(kw):
    from tools import diag_gemm as D
    assert D.conv_case(**kw)


def test_linear_rowmap_and_transposed_store():
    lib = _lib()
    from fgt_b200 import packing
    dev = torch.device("cuda:0")
    torch.manual_seed(1)
    M, N, K, Lb = 600, 256, 192, 150
    a = torch.randn(M, K, device=dev)
    w = torch.randn(N, K, device=dev) / K ** 0.5
    b = torch.randn(N, device=dev)
    ref = a.double() @ w.double().t() + b.double()
    a_s, w_s = lib.to_split(a), packing.pack_weight(w).to(dev)
    # row map: reverse order, drop every 7th row; residual add from the destination
    rowmap = torch.arange(M - 1, -1, -1, dtype=torch.int32, device=dev)
    rowmap[::7] = -1
    base = torch.randn(M, N, device=dev)
    out = base.clone()
    lib.gemm_tc([lib.ASeg(a_s, K, M)], w_s, N, out_w=M, bn=128, bias=b, out_f32=out, aux=out, aux_mode=lib.AUX_ADD,
                rowmap=rowmap)
    torch.cuda.synchronize()
    exp = base.double().clone()
    keep = rowmap >= 0
    exp[rowmap[keep].long()] += ref[keep]
    assert_close(out, exp, KTOL, "rowmap scatter + residual")
    # transposed, batched store: V^T[z, n, x] with pitch Lp
    Lp = 152
    vt = torch.zeros(2, M // Lb, N, Lp, dtype=torch.bfloat16, device=dev)
    lib.gemm_tc([lib.ASeg(a_s, K, M)], w_s, N, out_w=M, bn=128, bias=b, out_split=vt, lin_batch=Lb, os_z=N * Lp,
                os_x=1, os_c=Lp)
    torch.cuda.synchronize()
    got = lib.from_split(vt)[:, :, :Lb].permute(0, 2, 1).reshape(M, N)
    assert_close(got, ref, KTOL, "transposed store")


@pytest.mark.parametrize("batches,heads,L,qs", [(1, 1, 64, 1.0), (1, 2, 200, 1.0), (2, 4, 1800, 3.0), (4, 4, 37, 1.0)])
def test_attention_dense(batches, heads, L, qs):
    from tools import diag_attn as D
    assert D.dense_case(batches, heads, L, qscale=qs)


@pytest.mark.parametrize("frames,heads,nwin,nglob,qs", [(1, 1, 2, 60, 1.0), (3, 4, 15, 60, 3.0), (1, 4, 112, 448, 1.0)])
def test_attention_windowed(frames, heads, nwin, nglob, qs):
    from tools import diag_attn as D
    assert D.window_case(frames, heads, nwin, nglob, qscale=qs)


def test_rownorm_gather_affine():
    lib = _lib()
    dev = torch.device("cuda:0")
    torch.manual_seed(2)
    a = torch.randn(100, 512, device=dev) * 3 + 1
    b = torch.randn(100, 256, device=dev)
    gather = torch.randint(-1, 100, (2 * 70,), dtype=torch.int32, device=dev)
    out = torch.full((2, 2 * 80, 768), 7.0, dtype=torch.bfloat16, device=dev)
    gam, bet = torch.randn(768, device=dev), torch.randn(768, device=dev)
    lib.rownorm(a, b, out, gather=gather, rows_per_batch=70, total_rows=140, dst_batch_rows=80, dst_row0=5,
                gamma=gam, beta=bet)
    torch.cuda.synchronize()
    got = lib.from_split(out).reshape(2, 80, 768)[:, 5:75].reshape(140, 768)
    src = torch.cat([a, b], 1).double()
    ref = F.layer_norm(src[gather.clamp(min=0).long()], (768,), gam.double(), bet.double())
    ref[gather < 0] = 0
    assert_close(got, ref, KTOL, "rownorm")
    untouched = lib.from_split(out).reshape(2, 80, 768)[:, :5]
    assert (untouched == 14.0).all()


def test_fold_unfold_dwconv_dwpool_upsample_pack():
    lib = _lib()
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    bt, th, tw, C, OH, OW = 2, 6, 9, 40, 18, 25  # tokens of a 72x100 input
    hid = torch.randn(bt * th * tw, 49 * C, device=dev)
    # torch fold layout is channel-major (c*49+p); the kernels use position-major (p*C+c)
    cm = hid.reshape(bt, th * tw, 49, C).permute(0, 3, 2, 1).reshape(bt, C * 49, th * tw).double()
    img_ref = F.fold(cm, (OH, OW), (7, 7), stride=3, padding=3)
    cnt = F.fold(torch.ones_like(cm[:, :49]), (OH, OW), (7, 7), stride=3, padding=3)
    img = torch.empty(bt, OH, OW, C, device=dev)
    lib.fold(hid, bt, th, tw, C, 7, 7, 3, 3, OH, OW, normalize=True, out=img)
    torch.cuda.synchronize()
    assert_close(img.permute(0, 3, 1, 2), img_ref / cnt, KTOL, "fold normalised")
    add = torch.randn(bt, OH, OW, C, device=dev)
    sp = lib.empty_split((bt, OH, OW, C), dev)
    lib.fold(hid, bt, th, tw, C, 7, 7, 3, 3, OH, OW, normalize=False, add=add, out_split=sp)
    torch.cuda.synchronize()
    assert_close(lib.from_split(sp).permute(0, 3, 1, 2), img_ref + add.permute(0, 3, 1, 2).double(), KTOL, "fold+add")
    un = lib.empty_split((bt * th * tw, 49 * C), dev)
    lib.unfold(img, bt, th, tw, C, 7, 7, 3, 3, OH, OW, un, relu=True)
    torch.cuda.synchronize()
    ur = F.relu(F.unfold(img.permute(0, 3, 1, 2).double(), (7, 7), stride=3, padding=3))
    ur = ur.reshape(bt, C, 49, th * tw).permute(0, 3, 2, 1).reshape(bt * th * tw, 49 * C)
    assert_close(lib.from_split(un), ur, KTOL, "unfold+relu")
    # depthwise 3x3 + identity
    x = torch.randn(bt, th, tw, 64, device=dev)
    wt, bs = torch.randn(64, 1, 3, 3, device=dev), torch.randn(64, device=dev)
    o = torch.empty_like(x)
    osplit = lib.empty_split(x.shape, dev)
    lib.dwconv3x3_res(x, bt, th, tw, 64, wt.reshape(-1).contiguous(), bs, o, osplit)
    torch.cuda.synchronize()
    xr = x.permute(0, 3, 1, 2).double()
    ref = F.conv2d(xr, wt.double(), bs.double(), padding=1, groups=64) + xr
    assert_close(o.permute(0, 3, 1, 2), ref, KTOL, "dwconv3x3_res")
    assert_close(lib.from_split(osplit).permute(0, 3, 1, 2), ref, KTOL, "dwconv3x3_res split")
    # depthwise pooling over a zero-padded grid of [a ; b]
    a, b2 = torch.randn(bt, th, tw, 32, device=dev), torch.randn(bt, th, tw, 16, device=dev)
    wk, bk = torch.randn(48, 1, 4, 4, device=dev), torch.randn(48, device=dev)
    gh, gw = 2, 4  # padded grid 8 x 16
    po = torch.empty(bt, gh * gw, 48, device=dev)
    lib.dwpool(a, b2, bt, th, tw, 4, gh, gw, wk.reshape(-1).contiguous(), bk, po)
    torch.cuda.synchronize()
    cat = F.pad(torch.cat([a, b2], -1), (0, 0, 0, 16 - tw, 0, 8 - th)).permute(0, 3, 1, 2).double()
    pr = F.conv2d(cat, wk.double(), bk.double(), stride=4, groups=48).permute(0, 2, 3, 1).reshape(bt, gh * gw, 48)
    assert_close(po, pr, KTOL, "dwpool")
    # nearest x2
    s_in = lib.to_split(torch.randn(bt, 5, 7, 16, device=dev))
    s_out = lib.empty_split((bt, 10, 14, 16), dev)
    lib.upsample2x(s_in, bt, 5, 7, 16, s_out)
    torch.cuda.synchronize()
    ref = F.interpolate(lib.from_split(s_in).permute(0, 3, 1, 2), scale_factor=2).permute(0, 2, 3, 1)
    assert torch.equal(lib.from_split(s_out), ref)
    # NCHW pack with replication padding
    f0, f1 = torch.randn(bt, 3, 10, 12, device=dev), torch.randn(bt, 1, 10, 12, device=dev)
    pk = lib.empty_split((bt, 14, 16, 8), dev)
    lib.pack_nchw(f0, f1, pk, pad=2)
    torch.cuda.synchronize()
    ref = F.pad(torch.cat([f0, f1], 1), (2, 2, 2, 2), mode="replicate").permute(0, 2, 3, 1)
    got = lib.from_split(pk)
    assert_close(got[..., :4], ref, 1e-4, "pack_nchw")
    assert (got[..., 4:] == 0).all()


@pytest.mark.parametrize("n,H,W,cin,cout,act,nchw", [(2, 37, 53, 64, 3, 4, True), (3, 16, 20, 256, 2, 0, False),
                                                      (1, 8, 40, 128, 3, 2, True)])
def test_taps_as_n_conv(n, H, W, cin, cout, act, nchw):
    """Tiny-Cout 3x3 conv as a 1x1 GEMM over taps (pack_taps_as_n) + fgt_tapsum == F.conv2d (zero padding)."""
    lib = _lib()
    from fgt_b200 import packing
    dev = torch.device("cuda:0")
    torch.manual_seed(7)
    x = torch.randn(n, H, W, cin, device=dev)
    w = torch.randn(cout, cin, 3, 3, device=dev) / (9 * cin) ** 0.5
    b = torch.randn(cout, device=dev)
    ref = F.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), b.double(), padding=1)
    ref = {0: lambda t: t, 2: torch.relu, 4: torch.tanh}[act](ref)
    wt = packing.pack_weight(lib.pack_taps_as_n(w)).to(dev)
    y = torch.empty(32, n * H * W, device=dev)   # column-planar
    lib.gemm_tc([lib.ASeg(lib.to_split(x.reshape(n * H * W, cin)), cin, n * H * W)], wt, 32, out_w=n * H * W, bn=32,
                bias=torch.zeros(32, device=dev), out_f32=y, os_x=1, os_c=n * H * W)
    out = torch.empty((n, cout, H, W) if nchw else (n, H, W, cout), device=dev)
    lib.tapsum(y, n, H, W, cout, 3, b, act, out, nchw=nchw)
    got = out if nchw else out.permute(0, 3, 1, 2)
    assert_close(got, ref, KTOL, "taps-as-N conv")
