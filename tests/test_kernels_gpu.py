"""GPU: every kernel of libtbn_sm100.so, called through the C ABI, against the torch ops the
oracle is built from (computed on the CPU in fp32/fp64)."""
import math

import pytest
import torch
import torch.nn.functional as F

from cases import rel_err

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _s():
    return torch.cuda.current_stream().cuda_stream


def nhwc(x, dtype):  # (N,C,H,W) cpu -> NHWC device tensor
    return x.permute(0, 2, 3, 1).contiguous().to(DEV, dtype)


def from_nhwc(y):
    return y.float().cpu().permute(0, 3, 1, 2).contiguous()


def fold_w(w):  # (cout,cin,k,k) -> [k*k*cin][cout]
    cout, cin, k, _ = w.shape
    return w.permute(2, 3, 1, 0).reshape(k * k * cin, cout).contiguous()


def pack_tc(w, kpad=None):  # (cout,cin,k,k) -> bf16 [cout][k*k][cin_pad64]
    cout, cin, k, _ = w.shape
    t = w.permute(0, 2, 3, 1)
    if kpad is not None:
        K = k * k * cin
        out = torch.zeros(cout, (kpad + 63) // 64 * 64)
        out[:, :K] = t.reshape(cout, K)
        return out.to(torch.bfloat16)
    cp = (cin + 63) // 64 * 64
    out = torch.zeros(cout, k * k, cp)
    out[:, :, :cin] = t.reshape(cout, k * k, cin)
    return out.reshape(cout, -1).to(torch.bfloat16)


@pytest.mark.parametrize("n,cin,h,w,cout,k,stride,pad", [
    (2, 3, 37, 41, 64, 7, 2, 3),      # stem, odd sizes, scalar gather path
    (2, 1, 32, 45, 64, 7, 2, 3),      # audio stem
    (3, 64, 14, 14, 96, 3, 1, 1),
    (2, 96, 15, 13, 96, 3, 2, 1),     # strided 3x3
    (2, 192, 9, 9, 32, 1, 1, 0),
])
def test_conv_simt_fp32(lib, n, cin, h, w, cout, k, stride, pad):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)
    b = torch.randn(cout, generator=g)
    ref = F.relu(F.conv2d(x.double(), wt.double(), b.double(), stride=stride, padding=pad)).float()
    oh, ow = ref.shape[2:]
    xd, wd, bd = nhwc(x, torch.float32), fold_w(wt).to(DEV), b.to(DEV)
    y_cs, y_co = cout + 32, 16  # write into a channel slice of a wider buffer
    y = torch.full((n, oh, ow, y_cs), -7.0, device=DEV)
    lib.call("tbn_conv_simt", xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), y.data_ptr(), n, h, w, cin, cin, 0, cout,
             y_cs, y_co, k, stride, pad, 1, lib.TBN_F32, _s())
    torch.cuda.synchronize()
    got = from_nhwc(y[..., y_co:y_co + cout])
    assert rel_err(got, ref) < 1e-5
    assert (y[..., :y_co] == -7).all() and (y[..., y_co + cout:] == -7).all()  # neighbours untouched


def test_nchw_to_nhwc_and_im2col(lib):
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 3, 21, 19, generator=g)
    xd = x.to(DEV)
    y = torch.empty(2, 21, 19, 8, device=DEV, dtype=torch.bfloat16)
    lib.call("tbn_nchw_to_nhwc", xd.data_ptr(), y.data_ptr(), 2, 3, 21, 19, 8, lib.TBN_BF16, _s())
    torch.cuda.synchronize()
    assert torch.equal(y[..., :3].float().cpu(), x.permute(0, 2, 3, 1).to(torch.bfloat16).float())
    assert (y[..., 3:] == 0).all()
    # im2col straight from NCHW fp32 == unfold
    k, s, p = 7, 2, 3
    oh, ow = (21 + 6 - 7) // 2 + 1, (19 + 6 - 7) // 2 + 1
    K, kpad = 147, 160
    a = torch.empty(2 * oh * ow, kpad, device=DEV, dtype=torch.bfloat16)
    lib.call("tbn_im2col", xd.data_ptr(), a.data_ptr(), 2, 21, 19, 3, 3, 0, k, s, p, kpad, lib.TBN_F32, 1, _s())
    torch.cuda.synchronize()
    cols = F.unfold(x, k, padding=p, stride=s)  # (n, cin*k*k, L) ordered (c, r, s)
    cols = cols.view(2, 3, k * k, oh * ow).permute(0, 3, 2, 1).reshape(2 * oh * ow, K)  # -> (tap, c)
    assert torch.equal(a[:, :K].float().cpu(), cols.to(torch.bfloat16).float())
    assert (a[:, K:] == 0).all()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_pools(lib, dtype):
    g = torch.Generator().manual_seed(3)
    code = lib.TBN_F32 if dtype == torch.float32 else lib.TBN_BF16
    for (n, c, h, w) in ((2, 64, 21, 27), (1, 192, 14, 14), (2, 320, 8, 13), (1, 64, 112, 112)):
        x = torch.randn(n, c, h, w, generator=g).to(dtype).float()
        xd = nhwc(x, dtype)
        ref = F.max_pool2d(x, 3, 2, ceil_mode=True)
        oh, ow = ref.shape[2:]
        y = torch.zeros(n, oh, ow, c + 64, device=DEV, dtype=dtype)
        lib.call("tbn_maxpool3s2", xd.data_ptr(), y.data_ptr(), n, h, w, c, c, 0, c + 64, 64, code, _s())
        torch.cuda.synchronize()
        assert torch.equal(from_nhwc(y[..., 64:]), ref)
        for mode, fn in ((lib.POOL_AVG, lambda t: F.avg_pool2d(t, 3, 1, 1, ceil_mode=True, count_include_pad=True)),
                         (lib.POOL_MAX, lambda t: F.max_pool2d(t, 3, 1, 1, ceil_mode=True))):
            y2 = torch.empty(n, h, w, c, device=DEV, dtype=dtype)
            lib.call("tbn_pool3s1", xd.data_ptr(), y2.data_ptr(), n, h, w, c, c, 0, c, 0, mode, code, _s())
            torch.cuda.synchronize()
            tol = 1e-6 if dtype == torch.float32 else 8e-3
            assert rel_err(from_nhwc(y2), fn(x)) < tol
        # feature pooling
        if c % 8 == 0:
            f = torch.empty(n, c + 8, device=DEV)
            lib.call("tbn_avgpool_feat", xd.data_ptr(), f.data_ptr(), n, h, w, c, c + 8, 0, code, _s())
            fp = torch.empty(n * w, c, device=DEV)
            lib.call("tbn_avgpool_feat", xd.data_ptr(), fp.data_ptr(), n, h, w, c, c, 1, code, _s())
            torch.cuda.synchronize()
            assert rel_err(f[:, :c].cpu(), x.mean((2, 3))) < 1e-5
            assert rel_err(fp.cpu().view(n, w, c), x.mean(2).permute(0, 2, 1)) < 1e-5


def test_linear_and_grouped(lib):
    g = torch.Generator().manual_seed(4)
    for (m, n, k) in ((6, 477, 512), (192, 1024, 1024), (78, 1024, 1024), (5, 16, 256)):
        x = torch.randn(m, k + 8, generator=g)
        w = torch.randn(n, k, generator=g) / math.sqrt(k)
        b = torch.randn(3, n, generator=g)
        ref = F.relu(x[:, :k].double() @ w.double().t() + b.double()[torch.arange(m) % 3])
        y = torch.empty(m, n + 4, device=DEV)
        xd, wd, bd = x.to(DEV), w.to(DEV), b.to(DEV)  # keep the device copies alive across the launch
        lib.call("tbn_linear", xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), y.data_ptr(), m, n, k,
                 k + 8, k, n + 4, 3, 1, _s())
        torch.cuda.synchronize()
        assert rel_err(y[:, :n].cpu(), ref) < 1e-5
    # grouped: 4 heads, y[:, g*n:(g+1)*n] = x[:, g*k:(g+1)*k] @ w[g].T + b[g]
    m, n, k, G = 7, 64, 32, 4
    x = torch.randn(m, G * k, generator=g)
    w = torch.randn(G, n, k, generator=g)
    b = torch.randn(G, n, generator=g)
    ref = torch.cat([x[:, i * k:(i + 1) * k] @ w[i].t() + b[i] for i in range(G)], 1)
    y = torch.empty(m, G * n, device=DEV)
    xd, wd, bd = x.to(DEV), w.to(DEV), b.to(DEV)
    lib.call("tbn_linear_grouped", xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), y.data_ptr(), m, n,
             k, G * k, k, G * n, G, k, n * k, n, n, 0, _s())
    torch.cuda.synchronize()
    assert rel_err(y.cpu(), ref) < 1e-5


def test_groupnorm_mha_core_and_small_kernels(lib):
    g = torch.Generator().manual_seed(5)
    Fn, S, E, H = 5, 13, 1024, 4
    z = torch.randn(Fn, S, E, generator=g) * 2 + 0.3
    gamma, beta = torch.rand(E, generator=g) + 0.5, torch.randn(E, generator=g)
    ref = F.group_norm(z.permute(0, 2, 1), 64, gamma, beta, 1e-5).permute(0, 2, 1)
    zd = z.to(DEV).contiguous()
    gd, btd = gamma.to(DEV), beta.to(DEV)
    lib.call("tbn_groupnorm", zd.data_ptr(), gd.data_ptr(), btd.data_ptr(), Fn, S, E, 64, 1e-5, _s())
    torch.cuda.synchronize()
    assert rel_err(zd.cpu(), ref) < 1e-5

    qk = torch.randn(Fn, H, E, generator=g) / 8
    sc = torch.einsum("fhe,fpe->fhp", qk.double(), ref.double())
    a = torch.softmax(sc, -1)
    wref = a.mean(1)
    zbar_ref = torch.einsum("fhp,fpe->fhe", a, ref.double())
    p = wref + 1e-6
    p = p / p.sum(-1, keepdim=True)
    ent_ref = -(p * p.log()).sum(-1)
    zbar = torch.empty(Fn, H, E, device=DEV)
    wts = torch.empty(Fn, S, device=DEV)
    ent = torch.empty(Fn, device=DEV)
    qkd = qk.to(DEV)
    lib.call("tbn_mha_core", qkd.data_ptr(), zd.data_ptr(), zbar.data_ptr(), wts.data_ptr(), ent.data_ptr(), Fn, S,
             E, H, _s())
    torch.cuda.synchronize()
    assert rel_err(wts.cpu(), wref) < 1e-5
    assert rel_err(zbar.cpu(), zbar_ref) < 1e-5
    assert rel_err(ent.cpu(), ent_ref) < 1e-5
    assert torch.allclose(wts.sum(-1).cpu(), torch.ones(Fn), atol=1e-5)

    # fused variant: GroupNorm applied to the staged z inside the attention core, weighted value sums also as the
    # bf16 [hi | hi | lo] split operand -- the same bits as tbn_groupnorm + tbn_mha_core + tbn_split_bf16x3
    zraw = z.to(DEV).contiguous()
    zbar2, wts2, ent2 = torch.empty_like(zbar), torch.empty_like(wts), torch.empty_like(ent)
    zbar3 = torch.empty(Fn, H * 3 * E, device=DEV, dtype=torch.bfloat16)
    lib.call("tbn_mha_core_gn", qkd.data_ptr(), zraw.data_ptr(), gd.data_ptr(), btd.data_ptr(), 64, 1e-5, zbar2.data_ptr(),
             zbar3.data_ptr(), wts2.data_ptr(), ent2.data_ptr(), Fn, S, E, H, _s())
    split = torch.empty_like(zbar3)
    lib.call("tbn_split_bf16x3", zbar.data_ptr(), split.data_ptr(), Fn, H * E, H * E, E, _s())
    torch.cuda.synchronize()
    assert torch.equal(zbar2, zbar) and torch.equal(wts2, wts) and torch.equal(ent2, ent)
    assert torch.equal(zbar3, split)
    assert torch.equal(zraw.cpu(), z)  # the input is not normalised in place

    w = torch.softmax(torch.randn(Fn, S, generator=g), -1)
    y = torch.zeros(Fn, 2 * E, device=DEV)
    wd = w.to(DEV)
    lib.call("tbn_attn_weighted_sum", zd.data_ptr(), wd.data_ptr(), y.data_ptr() + E * 4, Fn, S, E, 2 * E, _s())
    torch.cuda.synchronize()
    assert rel_err(y[:, E:].cpu(), (zd.cpu() * w.unsqueeze(-1)).sum(1)) < 1e-5
    assert (y[:, :E] == 0).all()

    lg = torch.randn(9, 16, generator=g)
    sm = torch.empty(9, 16, device=DEV)
    lgd = lg.to(DEV)
    lib.call("tbn_softmax_rows", lgd.data_ptr(), sm.data_ptr(), 9, 16, _s())
    x = torch.randn(4 * 25, 477, generator=g)
    c = torch.empty(4, 477, device=DEV)
    xd = x.to(DEV)
    lib.call("tbn_consensus", xd.data_ptr(), c.data_ptr(), 4, 25, 477, _s())
    torch.cuda.synchronize()
    assert rel_err(sm.cpu(), torch.softmax(lg, -1)) < 1e-6
    assert rel_err(c.cpu(), x.view(4, 25, 477).mean(1)) < 1e-6


TC_CASES = [
    # n, cin, h, w, cout, k   (stride 1, pad k//2)
    (2, 64, 8, 8, 64, 1),        # single k-block, single tile
    (3, 192, 28, 28, 64, 1),     # 3a 1x1: flattened pixels, 3 k-blocks
    (2, 96, 14, 14, 128, 3),     # cin with a 32-channel tail, 3x3 halo via TMA OOB fill
    (5, 64, 7, 7, 96, 3),        # several images per TMA box
    (2, 64, 20, 35, 192, 3),     # non-square, partial boxes
    (4, 1056, 7, 7, 352, 1),     # wide K, N split in two tiles
    (1, 224, 9, 11, 224, 3),     # 4 k-chunks incl. tail, odd sizes
    (2, 256, 14, 14, 256, 3),    # N = 256 tile, 36 k-blocks (pipeline wrap-around)
    (3, 64, 56, 56, 192, 3),     # conv2_3x3: halo mode, single 64-channel slice, partial tiles in h
    (2, 96, 28, 28, 96, 3),      # halo mode with a 32-channel tail slice, partial tiles in w and h
    (1, 160, 64, 105, 96, 3),    # audio map: halo mode, odd width, 3 slices
    (40, 128, 28, 28, 160, 3),   # halo mode, several tiles per persistent CTA (ring + accumulator wrap-around)
    (300, 576, 14, 14, 512, 1),  # fused 1x1 shape: 2 N tiles, several tiles per persistent CTA
]


@pytest.mark.parametrize("n,cin,h,w,cout,k", TC_CASES)
@pytest.mark.parametrize("halo,cta2", [(1, 1), (0, 1), (1, 2), (0, 2), (1, 0), (1, 3), (0, 3)])  # cta2: 1 = pairs for wide N, 2 = always, 0 = never, 3 = default policy
def test_conv_tc_matches_conv2d(lib, n, cin, h, w, cout, k, halo, cta2):
    """tcgen05 implicit-GEMM conv vs F.conv2d on the same bf16-rounded operands (fp32 accumulate):
    only the accumulation order differs, so the result must agree to ~1e-5 before the bf16 store."""
    if halo == 0 and k == 1:
        pytest.skip("halo switch only affects 3x3 convs")
    lib.call("tbn_set_tuning", b"tc_halo", halo)
    lib.call("tbn_set_tuning", b"tc_cta2", cta2)
    g = torch.Generator().manual_seed(6)
    x = torch.randn(n, cin, h, w, generator=g).to(torch.bfloat16).float()
    wt = (torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)).to(torch.bfloat16).float()
    b = torch.randn(cout, generator=g)
    ref = F.relu(F.conv2d(x, wt, b, padding=k // 2))  # bf16-exact operands: fp32 conv is accurate to ~1e-6
    x_cs, x_co = cin + 64, 32  # read a channel slice of a wider buffer
    xd = torch.randn(n, h, w, x_cs, generator=g).to(DEV, torch.bfloat16)
    xd[..., x_co:x_co + cin] = nhwc(x, torch.bfloat16)
    wp, bd = pack_tc(wt).to(DEV), b.to(DEV)
    y_cs, y_co = cout + 48, 16
    # fp32 output: exact comparison of the accumulators
    y32 = torch.full((n, h, w, y_cs), -7.0, device=DEV)
    lib.call("tbn_conv_tc", xd.data_ptr(), wp.data_ptr(), bd.data_ptr(), y32.data_ptr(), n, h, w, cin, x_cs, x_co, cout,
             y_cs, y_co, k, 1, lib.TBN_F32, 1, _s())
    # bf16 output
    y16 = torch.full((n, h, w, y_cs), -7.0, device=DEV, dtype=torch.bfloat16)
    lib.call("tbn_conv_tc", xd.data_ptr(), wp.data_ptr(), bd.data_ptr(), y16.data_ptr(), n, h, w, cin, x_cs, x_co, cout,
             y_cs, y_co, k, 1, lib.TBN_BF16, 1, _s())
    torch.cuda.synchronize()
    assert rel_err(from_nhwc(y32[..., y_co:y_co + cout]), ref) < 1e-4
    assert (y32[..., :y_co] == -7).all() and (y32[..., y_co + cout:] == -7).all()
    assert rel_err(from_nhwc(y16[..., y_co:y_co + cout]), ref) < 5e-3
    assert (y16[..., :y_co] == -7).all() and (y16[..., y_co + cout:] == -7).all()
    lib.call("tbn_set_tuning", b"tc_halo", 1)
    lib.call("tbn_set_tuning", b"tc_cta2", 3)


@pytest.mark.parametrize("n,cin,h,w,cout", [
    (2, 128, 28, 28, 160),     # inception_3c_3x3
    (3, 96, 15, 13, 96),       # odd sizes, 32-channel tail slice
    (2, 256, 14, 14, 256),     # inception_4e_double_3x3_2
    (1, 128, 32, 52, 160),     # audio map
    (37, 192, 14, 14, 256),    # several tiles per persistent CTA
])
def test_conv_tc_stride2(lib, n, cin, h, w, cout):
    """3x3 / stride 2 / pad 1 through TMA boxes with element stride 2 (no patch matrix) vs F.conv2d."""
    g = torch.Generator().manual_seed(16)
    x = torch.randn(n, cin, h, w, generator=g).to(torch.bfloat16).float()
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(cin * 9)).to(torch.bfloat16).float()
    b = torch.randn(cout, generator=g)
    ref = F.relu(F.conv2d(x, wt, b, stride=2, padding=1))
    oh, ow = ref.shape[2:]
    x_cs, x_co = cin + 32, 16
    xd = torch.randn(n, h, w, x_cs, generator=g).to(DEV, torch.bfloat16)
    xd[..., x_co:x_co + cin] = nhwc(x, torch.bfloat16)
    wp, bd = pack_tc(wt).to(DEV), b.to(DEV)
    y_cs, y_co = cout + 64, 32
    y = torch.full((n, oh, ow, y_cs), -7.0, device=DEV, dtype=torch.bfloat16)
    lib.call("tbn_conv_tc_s2", xd.data_ptr(), wp.data_ptr(), bd.data_ptr(), y.data_ptr(), n, h, w, cin, x_cs, x_co, cout,
             y_cs, y_co, 1, lib.TBN_BF16, _s())
    torch.cuda.synchronize()
    assert rel_err(from_nhwc(y[..., y_co:y_co + cout]), ref) < 5e-3
    assert (y[..., :y_co] == -7).all() and (y[..., y_co + cout:] == -7).all()


@pytest.mark.parametrize("probs", [
    # (n, cin, h, w, cout, k, stride): the same layer of two modalities + the block's two independent 3x3 convs
    [(6, 64, 28, 28, 64, 3, 1), (6, 64, 28, 28, 96, 3, 1), (6, 64, 32, 52, 64, 3, 1), (6, 64, 32, 52, 96, 3, 1)],   # 3a {3x3, dbl_1}
    [(24, 192, 7, 7, 320, 3, 1), (24, 160, 7, 7, 224, 3, 1), (24, 192, 8, 13, 320, 3, 1), (24, 160, 8, 13, 224, 3, 1)],  # 5a (pairs)
    [(10, 576, 14, 14, 512, 1, 1), (10, 576, 16, 26, 512, 1, 1)],                                                     # 4a fused 1x1
    [(8, 128, 28, 28, 160, 3, 2), (8, 64, 28, 28, 96, 3, 1), (8, 128, 32, 52, 160, 3, 2), (8, 64, 32, 52, 96, 3, 1)],    # 3c: mixed variants
    [(3, 224, 7, 7, 224, 3, 1), (3, 1024, 7, 7, 128, 1, 1), (3, 224, 8, 13, 224, 3, 1), (3, 1024, 8, 13, 128, 1, 1),
     (3, 224, 7, 7, 224, 3, 1)],                                                                                        # 5b {dbl_2, pool_proj} x 3 modalities
    [(2, 96, 14, 14, 128, 3, 1)],                                                                                       # a group of one
])
def test_conv_tc_group_equals_separate_launches(lib, probs):
    """tbn_conv_tc_group (one persistent grid over the concatenated tile lists of several problems) must produce the
    same BITS as one tbn_conv_tc / tbn_conv_tc_s2 call per problem, and both must match F.conv2d."""
    g = torch.Generator().manual_seed(23)
    descs, refs, outs_g, outs_s, keep = [], [], [], [], []
    for (n, cin, h, w, cout, k, stride) in probs:
        x = torch.randn(n, cin, h, w, generator=g).to(torch.bfloat16).float()
        wt = (torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)).to(torch.bfloat16).float()
        b = torch.randn(cout, generator=g)
        ref = F.relu(F.conv2d(x, wt, b, stride=stride, padding=k // 2))
        oh, ow = ref.shape[2:]
        x_cs, x_co = cin + 32, 16
        xd = torch.randn(n, h, w, x_cs, generator=g).to(DEV, torch.bfloat16)
        xd[..., x_co:x_co + cin] = nhwc(x, torch.bfloat16)
        wp, bd = pack_tc(wt).to(DEV), b.to(DEV)
        y_cs, y_co = cout + 48, 16
        yg = torch.full((n, oh, ow, y_cs), -7.0, device=DEV, dtype=torch.bfloat16)
        ys = torch.full((n, oh, ow, y_cs), -7.0, device=DEV, dtype=torch.bfloat16)
        keep += [xd, wp, bd]
        descs.append(dict(x=xd.data_ptr(), wp=wp.data_ptr(), bias=bd.data_ptr(), segs=[(yg.data_ptr(), y_cs, y_co, cout)], n=n, h=h, w=w,
                          cin=cin, x_cs=x_cs, x_co=x_co, k=k, stride=stride, out_dtype=lib.TBN_BF16))
        if stride == 2:
            lib.call("tbn_conv_tc_s2", xd.data_ptr(), wp.data_ptr(), bd.data_ptr(), ys.data_ptr(), n, h, w, cin, x_cs, x_co, cout,
                     y_cs, y_co, 1, lib.TBN_BF16, _s())
        else:
            lib.call("tbn_conv_tc", xd.data_ptr(), wp.data_ptr(), bd.data_ptr(), ys.data_ptr(), n, h, w, cin, x_cs, x_co, cout,
                     y_cs, y_co, k, 1, lib.TBN_BF16, 1, _s())
        refs.append((ref, y_co, cout))
        outs_g.append(yg)
        outs_s.append(ys)
    arr, addr, ka = lib.make_descs(descs)
    lib.call("tbn_conv_tc_group", addr, len(descs), _s())
    torch.cuda.synchronize()
    for (ref, y_co, cout), yg, ys in zip(refs, outs_g, outs_s):
        assert rel_err(from_nhwc(yg[..., y_co:y_co + cout]), ref) < 5e-3
        assert (yg[..., :y_co] == -7).all() and (yg[..., y_co + cout:] == -7).all()
        assert torch.equal(yg, ys)


def test_conv_tc_strided_via_im2col_and_bias_table(lib):
    g = torch.Generator().manual_seed(7)
    n, cin, h, w, cout, k, s, p = 2, 3, 37, 41, 64, 7, 2, 3
    x = torch.randn(n, cin, h, w, generator=g)
    wt = (torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k))
    b = torch.randn(cout, generator=g)
    ref = F.relu(F.conv2d(x.to(torch.bfloat16).double(), wt.to(torch.bfloat16).double(), b.double(), stride=s, padding=p))
    oh, ow = ref.shape[2:]
    kpad = (k * k * cin + 15) // 16 * 16
    a = torch.empty(n * oh * ow, kpad, device=DEV, dtype=torch.bfloat16)
    xd, wpd, bd = x.to(DEV), pack_tc(wt, kpad).to(DEV), b.to(DEV)
    lib.call("tbn_im2col", xd.data_ptr(), a.data_ptr(), n, h, w, cin, cin, 0, k, s, p, kpad, lib.TBN_F32, 1, _s())
    y = torch.empty(n, oh, ow, cout, device=DEV, dtype=torch.bfloat16)
    lib.call("tbn_conv_tc", a.data_ptr(), wpd.data_ptr(), bd.data_ptr(), y.data_ptr(), n, oh,
             ow, kpad, kpad, 0, cout, cout, 0, 1, 1, lib.TBN_BF16, 1, _s())
    torch.cuda.synchronize()
    assert rel_err(from_nhwc(y), ref.float()) < 5e-3
    # GEMM with a per-position bias table (the PE-folded Conv1d of the attention module)
    m, kk, nn, S = 13 * 7, 128, 64, 13
    xa = torch.randn(m, kk, generator=g).to(torch.bfloat16)
    wa = (torch.randn(nn, kk, generator=g) / math.sqrt(kk)).to(torch.bfloat16)
    tb = torch.randn(S, nn, generator=g)
    ref2 = xa.double() @ wa.double().t() + tb.double()[torch.arange(m) % S]
    y2 = torch.empty(m, nn, device=DEV)
    xad, wad, tbd = xa.to(DEV), wa.to(DEV), tb.to(DEV)
    lib.call("tbn_conv_tc", xad.data_ptr(), wad.data_ptr(), tbd.data_ptr(), y2.data_ptr(), 1, 1, m, kk,
             kk, 0, nn, nn, 0, 1, 0, lib.TBN_F32, S, _s())
    torch.cuda.synchronize()
    assert rel_err(y2.cpu(), ref2) < 1e-4


def test_gemm_splitk_bf16x3(lib):
    """Head-GEMM path: fp32 activations split into [hi|hi|lo] bf16, weights [hi|lo|hi], split-K tcgen05 GEMM + reduce:
    fp32-class accuracy on a tall-K, tiny-M product."""
    g = torch.Generator().manual_seed(17)
    m, k, n, S = 26, 1024, 480, 13
    x = torch.randn(m, k + 64, generator=g)          # leading dimension larger than k
    w = torch.randn(n, k, generator=g) / math.sqrt(k)
    tb = torch.randn(S, n, generator=g)
    ref = F.relu(x[:, :k].double() @ w.double().t() + tb.double()[torch.arange(m) % S])
    hi = w.to(torch.bfloat16)
    lo = (w.double() - hi.double()).to(torch.bfloat16)
    w3 = torch.cat([hi, lo, hi], 1).contiguous().to(DEV)
    xd, tbd = x.to(DEV), tb.to(DEV)
    a3 = torch.empty(m, 3 * k, device=DEV, dtype=torch.bfloat16)
    lib.call("tbn_split_bf16x3", xd.data_ptr(), a3.data_ptr(), m, k, k + 64, k, _s())
    ksplit = 8
    part = torch.empty(ksplit, m, n, device=DEV)
    y = torch.full((m, n + 32), -7.0, device=DEV)
    lib.call("tbn_gemm_tc_splitk", a3.data_ptr(), w3.data_ptr(), part.data_ptr(), m, 3 * k, n, ksplit, _s())
    lib.call("tbn_splitk_reduce", part.data_ptr(), tbd.data_ptr(), y.data_ptr(), m, n, n + 32, ksplit, S, 1, _s())
    torch.cuda.synchronize()
    assert rel_err(y[:, :n].cpu(), ref) < 2e-5
    assert (y[:, n:] == -7).all()


def test_conv_tc_multi_segments(lib):
    """One fused GEMM for the three same-input 1x1 convs of an inception block, three destinations."""
    g = torch.Generator().manual_seed(8)
    n, cin, h, w = 3, 192, 14, 14
    widths = (64, 96, 32)
    x = torch.randn(n, cin, h, w, generator=g).to(torch.bfloat16).float()
    wt = (torch.randn(sum(widths), cin, 1, 1, generator=g) / math.sqrt(cin)).to(torch.bfloat16).float()
    b = torch.randn(sum(widths), generator=g)
    ref = F.relu(F.conv2d(x.double(), wt.double(), b.double())).float()
    xd, wpd, bd = nhwc(x, torch.bfloat16), pack_tc(wt).to(DEV), b.to(DEV)
    y0 = torch.full((n, h, w, 256), -7.0, device=DEV, dtype=torch.bfloat16)   # concat buffer, slot at 32
    y1 = torch.full((n, h, w, 96), -7.0, device=DEV, dtype=torch.bfloat16)    # dense temp
    y2 = torch.full((n, h, w, 64), -7.0, device=DEV, dtype=torch.bfloat16)    # temp, slot at 16
    # third segment: raw projection (no bias, no ReLU) as used by the commuted avg-pool branch
    arr, addr = lib.make_segs([(y0.data_ptr(), 256, 32, 64), (y1.data_ptr(), 96, 0, 96), (y2.data_ptr(), 64, 16, 32, 0)])
    lib.call("tbn_conv_tc_multi", xd.data_ptr(), wpd.data_ptr(), bd.data_ptr(), addr, 3, n, h, w, cin, cin, 0, 1,
             lib.TBN_BF16, _s())
    torch.cuda.synchronize()
    raw = F.conv2d(x.double(), wt.double()).float()
    assert rel_err(from_nhwc(y0[..., 32:96]), ref[:, :64]) < 5e-3
    assert rel_err(from_nhwc(y1), ref[:, 64:160]) < 5e-3
    assert rel_err(from_nhwc(y2[..., 16:48]), raw[:, 160:]) < 5e-3
    assert (y0[..., :32] == -7).all() and (y0[..., 96:] == -7).all()
    assert (y2[..., :16] == -7).all() and (y2[..., 48:] == -7).all()


@pytest.mark.parametrize("cin,n,h,w", [(3, 2, 37, 41), (1, 2, 64, 85), (3, 3, 224, 224), (1, 3, 256, 421), (3, 40, 224, 224),
                                       (10, 2, 37, 41), (10, 3, 224, 224), (10, 1, 31, 300)])
def test_stem_fused_conv1_pool1_conv2r(lib, cin, n, h, w):
    """conv1 7x7/2 + BN + ReLU -> MaxPool(3, 2, ceil) -> conv2_3x3_reduce + BN + ReLU in one kernel vs the torch ops with
    the same rounding points (fp16 operands and conv1 output inside the kernel, bf16 result)."""
    from attention_based_tbn_b200.plan import pack_stem_rows
    g = torch.Generator().manual_seed(19)
    x = torch.randn(n, cin, h, w, generator=g)
    w1 = torch.randn(64, cin, 7, 7, generator=g) / math.sqrt(49 * cin)
    b1 = torch.randn(64, generator=g)
    w2 = torch.randn(64, 64, 1, 1, generator=g) / 8
    b2 = torch.randn(64, generator=g)
    hf = lambda t: t.to(torch.float16).float()  # noqa: E731
    c1 = F.relu(F.conv2d(hf(x), hf(w1), b1, stride=2, padding=3))
    p1 = F.max_pool2d(hf(c1), 3, 2, ceil_mode=True)
    ref = F.relu(F.conv2d(p1, hf(w2), b2))
    ph, pw = ref.shape[2:]
    xd = x.to(DEV)
    w1p = pack_stem_rows(w1, b1, torch.float16).to(DEV)
    w2p = w2.reshape(64, 64).to(torch.float16).contiguous().to(DEV)
    b2d = b2.to(DEV)
    y = torch.full((n, ph, pw, 96), -7.0, device=DEV, dtype=torch.bfloat16)
    lib.call("tbn_stem_fused_tc", xd.data_ptr(), lib.TBN_F32, w1p.data_ptr(), w2p.data_ptr(), b2d.data_ptr(), y.data_ptr(), n, cin,
             h, w, 96, 16, _s())
    torch.cuda.synchronize()
    assert rel_err(from_nhwc(y[..., 16:80]), ref) < 6e-3
    assert (y[..., :16] == -7).all() and (y[..., 80:] == -7).all()
    # bf16 input tensor: converts exactly to the kernel's fp16 operand, so the output must be the same bits as for
    # the fp32 tensor holding the same (bf16-rounded) values
    xb = xd.to(torch.bfloat16)
    xr = xb.float()
    yb, yr = torch.full_like(y, -7.0), torch.full_like(y, -7.0)
    lib.call("tbn_stem_fused_tc", xb.data_ptr(), lib.TBN_BF16, w1p.data_ptr(), w2p.data_ptr(), b2d.data_ptr(), yb.data_ptr(), n, cin,
             h, w, 96, 16, _s())
    lib.call("tbn_stem_fused_tc", xr.data_ptr(), lib.TBN_F32, w1p.data_ptr(), w2p.data_ptr(), b2d.data_ptr(), yr.data_ptr(), n, cin,
             h, w, 96, 16, _s())
    torch.cuda.synchronize()
    assert torch.equal(yr, yb)
    assert rel_err(from_nhwc(yb[..., 16:80]), ref) < 2e-2


def test_avgpool_bias_relu_commutes_with_projection(lib):
    """relu(conv1x1(avgpool(x)) + b) == relu(avgpool(conv1x1_nobias(x)) + b): the commuted pool branch."""
    g = torch.Generator().manual_seed(10)
    n, cin, cout, h, w = 2, 64, 32, 9, 11
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, 1, 1, generator=g) / 8
    b = torch.randn(cout, generator=g)
    ref = F.relu(F.conv2d(F.avg_pool2d(x, 3, 1, 1, ceil_mode=True, count_include_pad=True), wt, b))
    pre = F.conv2d(x, wt)
    pd, bd = nhwc(pre, torch.float32), b.to(DEV)
    y = torch.zeros(n, h, w, cout + 8, device=DEV)
    lib.call("tbn_avgpool3s1_bias_relu", pd.data_ptr(), bd.data_ptr(), y.data_ptr(), n, h, w, cout, cout, 0, cout + 8, 8,
             lib.TBN_F32, _s())
    torch.cuda.synchronize()
    assert rel_err(from_nhwc(y[..., 8:]), ref) < 1e-5


@pytest.mark.parametrize("n,cin,h,w,cout", [(2, 64, 28, 28, 96), (3, 64, 56, 56, 64), (1, 64, 17, 23, 80), (5, 64, 56, 56, 192),
                                            (2, 64, 33, 40, 160)])
def test_conv_tc_resident_weights_mode(lib, n, cin, h, w, cout):
    """Opt-in mode kept as a measured negative result (tbn_set_tuning("tc_bres")): all nine taps of a <= 96-column
    weight tile stay in shared memory for the life of the CTA.  Must agree with the default mode bit for bit."""
    g = torch.Generator().manual_seed(77)
    x = torch.randn(n, cin, h, w, generator=g).to(torch.bfloat16).float()
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(cin * 9)).to(torch.bfloat16).float()
    b = torch.randn(cout, generator=g)
    xd, wp, bd = nhwc(x, torch.bfloat16), pack_tc(wt).to(DEV), b.to(DEV)
    outs = []
    for bres in (0, 1, 2):  # 2 = CTA pairs, each CTA keeps half of the (<= 192-column) weight tile
        lib.call("tbn_set_tuning", b"tc_bres", bres)
        y = torch.full((n, h, w, cout), -7.0, device=DEV)
        lib.call("tbn_conv_tc", xd.data_ptr(), wp.data_ptr(), bd.data_ptr(), y.data_ptr(), n, h, w, cin, cin, 0, cout, cout, 0, 3, 1,
                 lib.TBN_F32, 1, _s())
        torch.cuda.synchronize()
        outs.append(y)
    lib.call("tbn_set_tuning", b"tc_bres", 2)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    assert rel_err(from_nhwc(outs[1]), F.relu(F.conv2d(x, wt, b, padding=1))) < 1e-4


@pytest.mark.parametrize("knob,value,default", [("tc_pdl", 1, 0), ("tc_occ2", 0, 1), ("tc_max_bn", 128, 256), ("tc_max_bn", 64, 256)])
def test_conv_tc_tuning_knobs_do_not_change_results(lib, knob, value, default):
    """Tiling / launch knobs only move work between CTAs: every output element keeps its K order, so the bits agree."""
    g = torch.Generator().manual_seed(78)
    n, cin, h, w, cout = 3, 192, 14, 14, 256
    x = torch.randn(n, cin, h, w, generator=g).to(torch.bfloat16).float()
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(cin * 9)).to(torch.bfloat16).float()
    b = torch.randn(cout, generator=g)
    xd, wp, bd = nhwc(x, torch.bfloat16), pack_tc(wt).to(DEV), b.to(DEV)
    outs = []
    for v in (default, value):
        lib.call("tbn_set_tuning", knob.encode(), v)
        y = torch.full((n, h, w, cout), -7.0, device=DEV, dtype=torch.bfloat16)
        lib.call("tbn_conv_tc", xd.data_ptr(), wp.data_ptr(), bd.data_ptr(), y.data_ptr(), n, h, w, cin, cin, 0, cout, cout, 0, 3, 1,
                 lib.TBN_BF16, 1, _s())
        torch.cuda.synchronize()
        outs.append(y)
    lib.call("tbn_set_tuning", knob.encode(), default)
    assert torch.equal(outs[0], outs[1])
    assert rel_err(from_nhwc(outs[1]), F.relu(F.conv2d(x, wt, b, padding=1))) < 5e-3


def test_head_chain_fused_producers(lib):
    """The producers that write the head GEMMs' bf16 [hi | hi | lo] operands directly must give the same BITS as the
    separate launches they replace: feature pooling + split, split-K reduction (+ segment consensus) + split."""
    g = torch.Generator().manual_seed(31)
    # --- tbn_avgpool_feat_split (global and frequency-only pooling) == tbn_avgpool_feat + tbn_split_bf16x3
    for (n, h, w, freq) in ((5, 7, 7, 0), (3, 8, 13, 1)):
        c = 1024
        x = torch.randn(n, h, w, c, generator=g).to(DEV, torch.bfloat16)
        rows = n * (w if freq else 1)
        y0 = torch.zeros(rows, 2 * c, device=DEV)
        y1 = torch.zeros(rows, 2 * c, device=DEV)
        ld3 = 2 * 3 * c  # a block of a wider split buffer
        s0 = torch.zeros(rows, 3 * c, device=DEV, dtype=torch.bfloat16)
        s1 = torch.full((rows, ld3), 7.0, device=DEV, dtype=torch.bfloat16)
        lib.call("tbn_avgpool_feat", x.data_ptr(), y0.data_ptr() + c * 4, n, h, w, c, 2 * c, freq, lib.TBN_BF16, _s())
        lib.call("tbn_split_bf16x3", y0.data_ptr() + c * 4, s0.data_ptr(), rows, c, 2 * c, c, _s())
        lib.call("tbn_avgpool_feat_split", x.data_ptr(), y1.data_ptr() + c * 4, s1.data_ptr() + 3 * c * 2, ld3, n, h, w, c, 2 * c, freq,
                 lib.TBN_BF16, _s())
        torch.cuda.synchronize()
        assert torch.equal(y0, y1)
        assert torch.equal(s1[:, 3 * c:], s0) and (s1[:, :3 * c] == 7).all()
        s2 = torch.zeros(rows, 3 * c, device=DEV, dtype=torch.bfloat16)  # fp32 output is optional
        lib.call("tbn_avgpool_feat_split", x.data_ptr(), 0, s2.data_ptr(), 3 * c, n, h, w, c, 2 * c, freq, lib.TBN_BF16, _s())
        torch.cuda.synchronize()
        assert torch.equal(s2, s0)
    # --- tbn_splitk_reduce_ex == tbn_splitk_reduce (+ tbn_consensus) + tbn_split_bf16x3_ld
    ks, m, n, nseg = 4, 24, 512, 3
    part = torch.randn(ks, m, n, generator=g).to(DEV)
    bias = torch.randn(n, generator=g).to(DEV)
    full = torch.empty(m, n, device=DEV)
    lib.call("tbn_splitk_reduce", part.data_ptr(), bias.data_ptr(), full.data_ptr(), m, n, n, ks, 1, 1, _s())
    cons = torch.empty(m // nseg, n, device=DEV)
    lib.call("tbn_consensus", full.data_ptr(), cons.data_ptr(), m // nseg, nseg, n, _s())
    ld3 = 3 * n + 64
    sp_ref = torch.zeros(m // nseg, ld3, device=DEV, dtype=torch.bfloat16)
    lib.call("tbn_split_bf16x3_ld", cons.data_ptr(), sp_ref.data_ptr(), ld3, m // nseg, n, n, n, _s())
    y = torch.empty(m // nseg, n, device=DEV)
    sp = torch.zeros(m // nseg, ld3, device=DEV, dtype=torch.bfloat16)
    lib.call("tbn_splitk_reduce_ex", part.data_ptr(), bias.data_ptr(), y.data_ptr(), m, n, n, ks, 1, 1, nseg, sp.data_ptr(), ld3, _s())
    torch.cuda.synchronize()
    assert torch.equal(y, cons) and torch.equal(sp, sp_ref)
    ref = torch.relu(part.sum(0) + bias).view(m // nseg, nseg, n).mean(1)
    assert rel_err(y.cpu(), ref.cpu()) < 1e-5
    hi, lo = sp[:, :n].float(), sp[:, 2 * n:3 * n].float()
    assert torch.equal(sp[:, n:2 * n], sp[:, :n]) and rel_err((hi + lo).cpu(), ref.cpu()) < 1e-4
    # no consensus, fp32 only (nseg = 1, split = NULL): the plain reduction
    y1 = torch.empty(m, n, device=DEV)
    lib.call("tbn_splitk_reduce_ex", part.data_ptr(), bias.data_ptr(), y1.data_ptr(), m, n, n, ks, 1, 1, 1, 0, 0, _s())
    torch.cuda.synchronize()
    assert torch.equal(y1, full)
