"""Analytical per-SM cycle model of the backbone convs under different kernel strategies (design aid).
time(tile) = max(tensor cycles, shared-memory bytes / 128 B/clk, L2->SM bytes / L2BW B/clk/SM)."""
import math
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from attention_based_tbn_b200.arch import block_specs, spatial_trace, conv_out

SMS = 148
L2BW = float(os.environ.get("L2BW", 40))   # B/clk/SM delivered from L2
SMEMBW = 128.0
GHZ = 1.75


def layers(cin0, H, W, F):
    tr = spatial_trace(H, W)
    out = []
    h, w = tr["pool1"]
    out.append(("conv2r", F, h, w, 64, 64, 1, 1))
    out.append(("conv2", F, h, w, 64, 192, 3, 1))
    h, w = tr["pool2"]
    for b in block_specs():
        oh, ow = tr[b.name]
        n1 = (b.b1x1.cout if b.b1x1 else 0) + b.b3x3[0].cout + b.bdbl[0].cout + (b.bpool.cout if (b.bpool and b.pool == "avg") else 0)
        out.append((b.name + "_f1x1", F, h, w, b.cin, n1, 1, 1))
        if b.bpool and b.pool != "avg":
            out.append((b.name + "_pp", F, h, w, b.cin, b.bpool.cout, 1, 1))
        s = 2 if b.reduce else 1
        out.append((b.name + "_3x3", F, h, w, b.b3x3[1].cin, b.b3x3[1].cout, 3, s))
        out.append((b.name + "_d1", F, h, w, b.bdbl[1].cin, b.bdbl[1].cout, 3, 1))
        out.append((b.name + "_d2", F, h, w, b.bdbl[2].cin, b.bdbl[2].cout, 3, s))
        h, w = oh, ow
    return out


def pick_box(n, h, w):
    best = None
    for cw in range(1, min(w, 128) + 1):
        for ch in range(1, h + 1):
            if cw * ch > 128:
                break
            cn = min(128 // (cw * ch), n)
            t = math.ceil(w / cw) * math.ceil(h / ch) * math.ceil(n / cn)
            if best is None or t < best:
                best = t
    return best


def est(layer, strat):
    name, n, h, w, cin, cout, k, s = layer
    flops = 2.0 * n * conv_out(h, k, s, k // 2) * conv_out(w, k, s, k // 2) * cin * cout * k * k
    cta2 = "cta2" in strat
    if s == 2:  # im2col + GEMM: K = 9*cin as a 1x1 conv on the output pixels
        oh, ow = conv_out(h, 3, 2, 1), conv_out(w, 3, 2, 1)
        n, h, w, cin, k = 1, 1, n * oh * ow, 9 * cin, 1
    if k == 1:
        m_tiles = math.ceil(n * h * w / 128)
        halo = False
    else:
        m_tap = pick_box(n, h, w)
        m_halo = math.ceil(w / 8) * math.ceil(h / 16) * n
        halo = "halo" in strat and m_halo <= 1.15 * m_tap
        m_tiles = m_halo if halo else m_tap
    taps = k * k
    cch = math.ceil(cin / 64)
    nk16 = taps * (cin // 16)
    # N tile
    nt = math.ceil(cout / 256)
    bn = math.ceil(cout / nt / 16) * 16
    resident = False
    if "wres" in strat and halo:
        # weights resident if one N tile's weights fit ~112 KB (per CTA: half of them in cta2 mode)
        for cand_nt in range(nt, 5):
            cbn = math.ceil(cout / cand_nt / 16) * 16
            wbytes = taps * cch * 64 * cbn * 2 / (2 if cta2 else 1)
            if wbytes <= 112 * 1024 and cbn >= 64:
                nt, bn, resident = cand_nt, cbn, True
                break
    tensor = nk16 * bn / 2
    a_read = 4096
    b_read = 32 * bn / (2 if cta2 else 1)
    smem = nk16 * (a_read + b_read)
    bbytes = taps * cch * bn * 128 / (2 if cta2 else 1)
    if halo:
        abytes = cch * 23040
    else:
        abytes = taps * cch * 16384
    l2 = abytes + (0 if resident else bbytes)
    smem += l2 + 2 * 128 * bn * 2
    t = max(tensor, smem / SMEMBW, l2 / L2BW)
    tiles = m_tiles * nt
    rounds = math.ceil(tiles / SMS)
    cyc = rounds * t
    which = "T" if t == tensor else ("S" if t == smem / SMEMBW else "L")
    return flops, cyc, which, (m_tiles, nt, bn, halo, resident)


if __name__ == "__main__":
    strategies = ["tap", "halo", "halo+wres", "halo+cta2", "halo+wres+cta2"]
    total = {s: 0.0 for s in strategies}
    tf = 0.0
    for mod, (c0, H, W) in (("rgb", (3, 224, 224)), ("aud", (1, 256, 421))):
        for L in layers(c0, H, W, 192):
            row = f"{mod} {L[0]:10s} {L[2]:3d}x{L[3]:<3d} {L[4]:4d}->{L[5]:<4d} k{L[6]}s{L[7]}"
            for s in strategies:
                fl, cyc, which, info = est(L, s)
                total[s] += cyc
                row += f" | {cyc/1e3:7.1f}k {which} {fl/cyc/1e3*SMS*GHZ/1e3:5.0f}"
            tf += fl
            if "-v" in sys.argv:
                print(row, info)
    print("strategy: total cycles/SM, us @%.2f GHz, TF/s" % GHZ)
    for s in strategies:
        us = total[s] / GHZ / 1e3
        print(f"  {s:18s} {total[s]/1e6:7.2f} Mcyc  {us:8.1f} us   {tf/us/1e6:7.1f} TF/s")
