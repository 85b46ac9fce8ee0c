"""BN-Inception layer table used by the TBN backbones (shape contract of the conv kernels).

The reference takes the backbone from the third-party ``pretrainedmodels`` package
(reference core/models/bn_inception.py:5-11); the only in-tree statement of the architecture
is the vendored layer table in reference core/models/bn_inception_audio.py:21-39 (stem),
58-401 (inception 3a..5b) and its dataflow 437-1003.  This module restates that table as
data (one row per inception block) instead of ~1000 lines of module definitions; every
consumer (parameter container, execution plan, FLOP accounting) is generated from it.

Block row: (name, cin, n1x1, n3x3_reduce, n3x3, ndbl_reduce, ndbl_1, ndbl_2, pool, npool_proj)
  * normal blocks : concat order [1x1, 3x3, double_3x3_2, pool_proj]   (ref :485-493)
  * reduce blocks : concat order [3x3(s2), double_3x3_2(s2), maxpool(s2, ceil)] (ref :593-600)
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterator, List, Optional, Tuple

FEATURE_SIZE = 1024  # reference bn_inception.py:99

# name, cin, 1x1, 3x3r, 3x3, d3r, d3_1, d3_2, pool kind, pool_proj
INCEPTION_BLOCKS: Tuple[tuple, ...] = (
    ("3a", 192, 64, 64, 64, 64, 96, 96, "avg", 32),
    ("3b", 256, 64, 64, 96, 64, 96, 96, "avg", 64),
    ("3c", 320, 0, 128, 160, 64, 96, 96, "max_s2", 0),
    ("4a", 576, 224, 64, 96, 96, 128, 128, "avg", 128),
    ("4b", 576, 192, 96, 128, 96, 128, 128, "avg", 128),
    ("4c", 576, 160, 128, 160, 128, 160, 160, "avg", 128),
    ("4d", 608, 96, 128, 192, 160, 192, 192, "avg", 128),
    ("4e", 608, 0, 128, 192, 192, 256, 256, "max_s2", 0),
    ("5a", 1056, 352, 192, 320, 160, 224, 224, "avg", 128),
    ("5b", 1024, 352, 192, 320, 192, 224, 224, "max_s1", 128),
)


@dataclass(frozen=True)
class ConvSpec:
    """One conv + BatchNorm(eps=1e-5) + ReLU triple (Cadene naming: <name>, <name>_bn)."""

    name: str
    cin: int
    cout: int
    k: int
    stride: int
    pad: int


@dataclass(frozen=True)
class BlockSpec:
    name: str
    cin: int
    cout: int
    reduce: bool
    pool: str  # "avg" | "max_s1" | "max_s2"
    b1x1: Optional[ConvSpec]
    b3x3: Tuple[ConvSpec, ConvSpec]
    bdbl: Tuple[ConvSpec, ConvSpec, ConvSpec]
    bpool: Optional[ConvSpec]

    @property
    def out_offsets(self) -> dict:
        """Channel offset of every branch inside the block's concat buffer."""
        off, res = 0, {}
        for key, width in (
            ("1x1", self.b1x1.cout if self.b1x1 else 0),
            ("3x3", self.b3x3[1].cout),
            ("dbl", self.bdbl[2].cout),
            ("pool", self.bpool.cout if self.bpool else (self.cin if self.reduce else 0)),
        ):
            res[key] = off
            off += width
        assert off == self.cout, (self.name, off, self.cout)
        return res


def stem_specs(in_channels: int) -> List[ConvSpec]:
    """conv1_7x7_s2 -> pool1 -> conv2_3x3_reduce -> conv2_3x3 -> pool2 (ref :21-39, 417-426)."""
    return [
        ConvSpec("conv1_7x7_s2", in_channels, 64, 7, 2, 3),
        ConvSpec("conv2_3x3_reduce", 64, 64, 1, 1, 0),
        ConvSpec("conv2_3x3", 64, 192, 3, 1, 1),
    ]


def block_specs() -> List[BlockSpec]:
    out = []
    for (nm, cin, n1, n3r, n3, d3r, d31, d32, pool, npp) in INCEPTION_BLOCKS:
        p = f"inception_{nm}"
        reduce = pool == "max_s2"
        s = 2 if reduce else 1
        b1 = ConvSpec(f"{p}_1x1", cin, n1, 1, 1, 0) if n1 else None
        b3 = (ConvSpec(f"{p}_3x3_reduce", cin, n3r, 1, 1, 0), ConvSpec(f"{p}_3x3", n3r, n3, 3, s, 1))
        bd = (
            ConvSpec(f"{p}_double_3x3_reduce", cin, d3r, 1, 1, 0),
            ConvSpec(f"{p}_double_3x3_1", d3r, d31, 3, 1, 1),
            ConvSpec(f"{p}_double_3x3_2", d31, d32, 3, s, 1),
        )
        bp = ConvSpec(f"{p}_pool_proj", cin, npp, 1, 1, 0) if npp else None
        cout = n1 + n3 + d32 + (npp if npp else cin)
        out.append(BlockSpec(nm, cin, cout, reduce, pool, b1, b3, bd, bp))
    # chain consistency: every block consumes the previous block's concat width
    prev = 192
    for b in out:
        assert b.cin == prev, (b.name, b.cin, prev)
        prev = b.cout
    assert prev == FEATURE_SIZE
    return out


def all_conv_specs(in_channels: int) -> Iterator[ConvSpec]:
    """All 69 conv-BN-ReLU triples in Cadene definition order."""
    yield from stem_specs(in_channels)
    for b in block_specs():
        if b.b1x1:
            yield b.b1x1
        yield from b.b3x3
        yield from b.bdbl
        if b.bpool:
            yield b.bpool


def conv_out(n: int, k: int, s: int, p: int) -> int:
    return (n + 2 * p - k) // s + 1


def pool_out_ceil(n: int, k: int = 3, s: int = 2) -> int:
    """MaxPool2d(3, stride 2, ceil_mode=True, no padding): last window must start inside."""
    o = -((n - k) // -s) + 1
    if (o - 1) * s >= n:
        o -= 1
    return o


def spatial_trace(h: int, w: int) -> dict:
    """Output H,W after each stage for an input of (h, w)."""
    t = {}
    h, w = conv_out(h, 7, 2, 3), conv_out(w, 7, 2, 3)
    t["conv1"] = (h, w)
    h, w = pool_out_ceil(h), pool_out_ceil(w)
    t["pool1"] = (h, w)
    t["conv2"] = (h, w)
    h, w = pool_out_ceil(h), pool_out_ceil(w)
    t["pool2"] = (h, w)
    for b in block_specs():
        if b.reduce:
            h2, w2 = conv_out(h, 3, 2, 1), conv_out(w, 3, 2, 1)
            assert (h2, w2) == (pool_out_ceil(h), pool_out_ceil(w)), (b.name, h, w)
            h, w = h2, w2
        t[b.name] = (h, w)
    return t


def backbone_flops(in_channels: int, h: int, w: int) -> int:
    """2*MAC over the 69 convs for one frame (SURVEY.md section 8d figures)."""
    tr = spatial_trace(h, w)
    total = 0
    st = stem_specs(in_channels)
    for spec, key in zip(st, ("conv1", "conv2", "conv2")):
        oh, ow = tr[key]
        total += 2 * oh * ow * spec.cout * spec.cin * spec.k * spec.k
    prev = tr["pool2"]
    for b in block_specs():
        oh, ow = tr[b.name]
        ih, iw = prev
        for c in ([b.b1x1] if b.b1x1 else []) + [b.b3x3[0], b.bdbl[0], b.bdbl[1]]:
            total += 2 * ih * iw * c.cout * c.cin * c.k * c.k
        for c in (b.b3x3[1], b.bdbl[2]):
            total += 2 * oh * ow * c.cout * c.cin * c.k * c.k
        if b.bpool:
            total += 2 * oh * ow * b.bpool.cout * b.bpool.cin
        prev = (oh, ow)
    return total
