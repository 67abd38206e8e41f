"""Minimal YOLOv8 detection network, random-init, written from the published yolov8.yaml topology:

    backbone  0 Conv /2   1 Conv /4   2 C2f   3 Conv /8   4 C2f (C3)   5 Conv /16   6 C2f (C4)
              7 Conv /32  8 C2f       9 SPPF (C5)
    neck/head 10 Upsample 11 Concat[-1,6] 12 C2f 13 Upsample 14 Concat[-1,4] 15 C2f 16 Conv /2
              17 Concat[-1,12] 18 C2f 19 Conv /2 20 Concat[-1,9] 21 C2f 22 Detect[15,18,21]

Scales: n = (depth 0.33, width 0.25), s = (0.33, 0.50).  Every layer carries the `.i` / `.f` (index / from)
attributes and the class names (`SPPF`, `Concat`, `Detect`, ...) the reference's
`MCAQYOLO._find_backbone_out_indices` (models/mcaq_yolo.py:351-400) inspects, so that discovery yields
[4, 6, 9] with channels 64/128/256 (n) or 128/256/512 (s).  No checkpoints, no decode / NMS: eval mode returns
(concatenated raw maps, raw list) in the shape convention `_extract_raw_maps` (models/mcaq_yolo.py:21-38) expects.
"""
import math

import torch
import torch.nn as nn


class Conv(nn.Module):
    def __init__(self, c1, c2, k=1, s=1):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, k, s, k // 2, bias=False)
        self.bn = nn.BatchNorm2d(c2, eps=1e-3, momentum=0.03)
        self.act = nn.SiLU(inplace=True)

    def forward(self, x):
        return self.act(self.bn(self.conv(x)))


class Bottleneck(nn.Module):
    def __init__(self, c, shortcut=True):
        super().__init__()
        self.cv1 = Conv(c, c, 3)
        self.cv2 = Conv(c, c, 3)
        self.add = shortcut

    def forward(self, x):
        y = self.cv2(self.cv1(x))
        return x + y if self.add else y


class C2f(nn.Module):
    def __init__(self, c1, c2, n=1, shortcut=False):
        super().__init__()
        self.c = c2 // 2
        self.cv1 = Conv(c1, 2 * self.c, 1)
        self.cv2 = Conv((2 + n) * self.c, c2, 1)
        self.m = nn.ModuleList(Bottleneck(self.c, shortcut) for _ in range(n))

    def forward(self, x):
        y = list(self.cv1(x).chunk(2, 1))
        for m in self.m:
            y.append(m(y[-1]))
        return self.cv2(torch.cat(y, 1))


class SPPF(nn.Module):
    def __init__(self, c1, c2, k=5):
        super().__init__()
        c_ = c1 // 2
        self.cv1 = Conv(c1, c_, 1)
        self.cv2 = Conv(c_ * 4, c2, 1)
        self.m = nn.MaxPool2d(k, 1, k // 2)

    def forward(self, x):
        y = [self.cv1(x)]
        for _ in range(3):
            y.append(self.m(y[-1]))
        return self.cv2(torch.cat(y, 1))


class Concat(nn.Module):
    def forward(self, xs):
        return torch.cat(xs, 1)


class Detect(nn.Module):
    def __init__(self, nc, ch):
        super().__init__()
        self.nc, self.reg_max, self.nl = nc, 16, len(ch)
        self.no = nc + 4 * self.reg_max
        c2, c3 = max(16, ch[0] // 4, 4 * self.reg_max), max(ch[0], min(nc, 100))
        self.cv2 = nn.ModuleList(nn.Sequential(Conv(x, c2, 3), Conv(c2, c2, 3), nn.Conv2d(c2, 4 * self.reg_max, 1)) for x in ch)
        self.cv3 = nn.ModuleList(nn.Sequential(Conv(x, c3, 3), Conv(c3, c3, 3), nn.Conv2d(c3, nc, 1)) for x in ch)
        self.stride = torch.tensor([8.0, 16.0, 32.0])

    def forward(self, xs):
        raw = [torch.cat((self.cv2[i](x), self.cv3[i](x)), 1) for i, x in enumerate(xs)]
        if self.training:
            return raw
        return torch.cat([r.flatten(2) for r in raw], 2), raw


SCALES = {"n": (0.33, 0.25), "s": (0.33, 0.50), "m": (0.67, 0.75)}


class DetectionModel(nn.Module):
    """`.model` is the nn.Sequential the reference indexes (`model.model[idx]`, models/mcaq_yolo.py:459-473)."""

    def __init__(self, scale="n", nc=80, seed=0):
        super().__init__()
        d, w = SCALES[scale]
        ch = lambda c: int(math.ceil(min(c, 1024) * w / 8) * 8)      # noqa: E731
        n = lambda k: max(round(k * d), 1)                           # noqa: E731
        torch.manual_seed(seed)
        c64, c128, c256, c512, c1024 = ch(64), ch(128), ch(256), ch(512), ch(1024)
        spec = [
            (-1, Conv(3, c64, 3, 2)), (-1, Conv(c64, c128, 3, 2)), (-1, C2f(c128, c128, n(3), True)),
            (-1, Conv(c128, c256, 3, 2)), (-1, C2f(c256, c256, n(6), True)),
            (-1, Conv(c256, c512, 3, 2)), (-1, C2f(c512, c512, n(6), True)),
            (-1, Conv(c512, c1024, 3, 2)), (-1, C2f(c1024, c1024, n(3), True)), (-1, SPPF(c1024, c1024, 5)),
            (-1, nn.Upsample(None, 2, "nearest")), ([-1, 6], Concat()), (-1, C2f(c1024 + c512, c512, n(3))),
            (-1, nn.Upsample(None, 2, "nearest")), ([-1, 4], Concat()), (-1, C2f(c512 + c256, c256, n(3))),
            (-1, Conv(c256, c256, 3, 2)), ([-1, 12], Concat()), (-1, C2f(c256 + c512, c512, n(3))),
            (-1, Conv(c512, c512, 3, 2)), ([-1, 9], Concat()), (-1, C2f(c512 + c1024, c1024, n(3))),
            ([15, 18, 21], Detect(nc, (c256, c512, c1024))),
        ]
        layers = []
        for i, (f, m) in enumerate(spec):
            m.i, m.f = i, f
            layers.append(m)
        self.model = nn.Sequential(*layers)
        self.save = sorted({j for f, _ in spec for j in (f if isinstance(f, list) else [f]) if j != -1})
        self.args = None
        self.nc = nc
        self.stride = torch.tensor([8.0, 16.0, 32.0])
        self.names = {i: str(i) for i in range(nc)}
        self.yaml = {"nc": nc, "scale": scale}

    def forward(self, x):
        y = []
        for m in self.model:
            if m.f != -1:
                x = y[m.f] if isinstance(m.f, int) else [x if j == -1 else y[j] for j in m.f]
            x = m(x)
            y.append(x if m.i in self.save else None)
        return x
