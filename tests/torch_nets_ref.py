"""Plain torch (autograd) restatements of the three tile-level networks in training mode, with the reference's
semantics (core/morphology.py:81-97, 309-354, 962-968; core/bit_allocation.py:218-280; core/quantization.py:213-239):
the comparator of the native training kernels (csrc/train_nets.cu) in tests/test_gpu_train_nets.py."""
import torch
import torch.nn.functional as F


def bilateral(cmap, sigma_spatial=2.0, sigma_range=0.1, k=5):
    B, H, W = cmap.shape
    r = k // 2
    nb = F.unfold(F.pad(cmap.unsqueeze(1), (r, r, r, r), mode="replicate"), k)
    centre = cmap.reshape(B, 1, H * W)
    ax = torch.arange(k, device=cmap.device, dtype=torch.float32) - r
    d2 = ax.view(-1, 1) ** 2 + ax.view(1, -1) ** 2
    ws = torch.exp(-d2 / (2 * sigma_spatial ** 2)).reshape(1, -1, 1)
    wgt = ws * torch.exp(-((nb - centre) ** 2) / (2 * sigma_range ** 2))
    return ((wgt * nb).sum(1) / (wgt.sum(1) + 1e-8)).reshape(B, H, W)


def complexity(analyzer, phi):
    B, ht, wt, _ = phi.shape
    c = analyzer.complexity_mlp(phi.reshape(-1, 8)).reshape(B, ht, wt)
    return bilateral(c).clamp(0.0, 1.0)


def mapper(m, c, temperature, continuous=True):
    c = c.clamp(0.0, 1.0)
    B, H, W = c.shape
    x = c.reshape(-1, 1)
    h = m.mapping_network(torch.cat([x, x ** 2, torch.log1p(x)], dim=-1))
    bits = (m.min_bits + (m.max_bits - m.min_bits) * h).reshape(B, H, W)
    if temperature is not None:
        bits = bits * max(float(temperature), 0.1)
    bits = bits + (bits.clamp(m.min_bits, m.max_bits) - bits).detach()
    if not continuous:
        bits = bits + (torch.round(bits) - bits).detach()
    return bits


def soft_mask(sm, bit_map, x):
    B, C, H, W = x.shape
    Ht, Wt = bit_map.shape[-2:]
    with torch.no_grad():
        act = F.adaptive_avg_pool2d(x.abs().mean(1, keepdim=True).float(), (Ht, Wt))
        act = act / (act.amax(dim=(2, 3), keepdim=True) + 1e-8)
    bn = ((bit_map.unsqueeze(1).float() - 2.0) / 6.0).clamp(0.0, 1.0)
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        m = torch.softmax(sm.net(torch.cat([bn, act], 1)), dim=1)[:, :1]
        m = F.interpolate(m, size=(H, W), mode="nearest")
        p = sm.kernel_size // 2
        return F.conv2d(F.pad(m, (p, p, p, p), mode="replicate"), sm.smooth_kernel)
