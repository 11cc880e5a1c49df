"""Launch helpers shared by the LAFC / RAFT host modules: NHWC split-bf16 convolutions on the
tensor-core engine (fgt_gemm_tc), with the tile heuristics in one place."""
from . import lib
from .packing import pack_weight, pack_weight_im2col


def pick_box(ow, oh):
    """Output tile (box_w x box_h <= 128 positions) minimising padded work for an ow x oh map."""
    best = None
    for bw, bh in ((16, 8), (32, 4), (8, 16), (64, 2), (128, 1)):
        tiles = -(-ow // bw) * -(-oh // bh)
        if best is None or tiles < best[0]:
            best = (tiles, bw, bh)
    return best[1], best[2]


def pick_bn(n_per_group, groups=1):
    if n_per_group % 128 == 0:
        return 128
    if n_per_group % 96 == 0 and n_per_group <= 384:
        return 96
    if n_per_group <= 128:
        return (n_per_group + 15) // 16 * 16 if groups == 1 else n_per_group
    if groups > 1:
        for bn in (96, 80, 64, 48, 32, 16):
            if n_per_group % bn == 0:
                return bn
    return 128


def fill_sms(bn, n_total, m_positions, groups=1, n_sms=148):
    """Small-M problems (RAFT's 1/8-resolution update block: 51 tiles of 128 positions) leave most SMs idle
    with 128-wide channel tiles; halve the tile while that raises the number of busy SMs."""
    m_tiles = -(-m_positions // 128)
    while bn >= 128 and bn % 32 == 0 and m_tiles * -(-n_total // bn) < n_sms and (groups == 1 or (n_total // groups) % (bn // 2) == 0):
        bn //= 2
    return bn


def packed(name, weight, bias, dev, seg_counts=None, im2col_pad=None):
    """Weight record consumed by conv()/linear(): split-bf16 packed weight + fp32 bias."""
    if im2col_pad is not None:
        w = pack_weight_im2col(weight, im2col_pad)
    else:
        w = pack_weight(weight, seg_counts)
    return dict(w=w.to(dev), b=bias.detach().float().contiguous().to(dev), N=weight.shape[0], name=name)


def conv(xs, wp, *, kx=1, ky=1, kz=1, stride=1, dil=1, pad_x=0, pad_y=0, pad_z=0, out_z=None, act=lib.ACT_LEAKY02,
         out_split=None, out_f32=None, aux=None, aux_mode=lib.AUX_NONE, groups=1, seg_counts=None, nchw_out=False,
         alpha=1.0, aux2=None, out_c_total=None, out_c_offset=0):
    """xs: list of (split tensor [2, Z, Y, X, C], C) sharing Z/Y/X. Output NHWC [Z', Y', X', N]
    (or NCHW fp32 when nchw_out). Returns (oz, oy, ox)."""
    x0, c0 = xs[0]
    Z, Y, X = x0.shape[1], x0.shape[2], x0.shape[3]
    oy = (Y + 2 * pad_y - dil * (ky - 1) - 1) // stride + 1
    ox = (X + 2 * pad_x - dil * (kx - 1) - 1) // stride + 1
    oz = (Z + 2 * pad_z - (kz - 1)) if out_z is None else out_z
    N = wp["N"]
    segs = []
    for i, (t, c) in enumerate(xs):
        cnt = c if seg_counts is None else seg_counts[i]
        segs.append(lib.ASeg(t, c, X, Y, Z, c_per_group=(cnt if groups > 1 else 0), c_count=cnt))
    bw, bh = pick_box(ox, oy)
    if nchw_out:
        strides = dict(os_z=N * oy * ox, os_y=ox, os_x=1, os_c=oy * ox)
    else:
        ct = N if out_c_total is None else out_c_total  # write into a channel slice of a wider NHWC buffer
        strides = dict(os_z=oy * ox * ct, os_y=ox * ct, os_x=ct, os_c=1, out_elem_offset=out_c_offset)
    lib.gemm_tc(segs, wp["w"], N, kx=kx, ky=ky, kz=kz, stride=stride, dil=dil, pad_x=pad_x, pad_y=pad_y, pad_z=pad_z,
                groups=groups, out_w=ox, out_h=oy, out_z=oz, box_w=bw, box_h=bh,
                bn=fill_sms(pick_bn(N // groups, groups), N, ox * oy * oz, groups),
                bias=wp["b"], alpha=alpha, act=act, aux=aux, aux_mode=aux_mode, aux2=aux2, out_f32=out_f32, out_split=out_split,
                tag=wp["name"], **strides)
    return oz, oy, ox


def linear(segs, wp, rows, **kw):
    """segs: list of (split tensor [2, rows, C], C). Row-major [rows, N] outputs."""
    asegs = [lib.ASeg(t, c, rows) for t, c in segs]
    lib.gemm_tc(asegs, wp["w"], wp["N"], out_w=rows, bn=fill_sms(pick_bn(wp["N"], 1), wp["N"], rows),
                bias=wp["b"], tag=wp["name"], **kw)
