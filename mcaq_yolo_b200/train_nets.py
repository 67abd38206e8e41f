"""Training forms of the three tile-level networks as autograd Functions over libmcaq_b200.so
(csrc/train_nets.cu): forward with saved statistics, backward in one or two launches each, instead of the
~10^3 eager autograd kernels the reference's torch modules issue per step.

Each Function takes the network's parameters as ONE flat tensor (`flat_params`: torch.cat of the parameters in
module order, so autograd routes the flat gradient back to every nn.Parameter) next to the inputs.
CUDA tensors only; nothing synchronises (graph-capturable)."""
from __future__ import annotations

import ctypes
import weakref

import torch

from . import _lib, ops
from . import constants as K

CMLP_PARAMS, MAPPER_PARAMS, SOFTMASK_PARAMS = 2881, 4609, 170


_FLAT_CACHE = weakref.WeakKeyDictionary()


def flat_is_cuda(t) -> bool:
    return t.is_cuda


def prepare_step(*owners):
    """Build the shared parameter concatenations of `owners` (modules passed to flat_params(owner=...)) on the CURRENT
    stream.  Call it on the main stream before the scales fork onto their own streams: autograd accumulates the
    gradient of a shared node on the stream that created it, so a node created inside the first scale's stream would
    make that scale's whole backward wait for the other scales' contributions."""
    for o in owners:
        flat_params(o.parameters(), owner=o)
_COUNTS = {}


def flat_params(params, owner=None) -> torch.Tensor:
    """torch.cat of the (flattened) parameters: differentiable, gradients flow back to each parameter.

    With `owner` (the module holding the parameters) the concatenation is shared by every forward of one training
    step: the three scales of a step call the same nets, and one shared node means autograd sums their three flat
    gradients (two adds) and splits ONCE, instead of accumulating 3 x P per-parameter slices (the 83 small adds of
    profiles/r02_train_step_launches.txt).  The shared node is dropped as soon as a backward pass reaches it, or
    when any parameter changes (optimizer step: `_version`; re-assignment: data pointer)."""
    params = list(params)
    if owner is None or not torch.is_grad_enabled():
        return torch.cat([p.reshape(-1).float() for p in params])
    capturing = bool(params) and params[0].is_cuda and torch.cuda.is_current_stream_capturing()
    # (a node built inside a graph capture is never reused outside it, nor the other way round)
    key = tuple((p.data_ptr(), p._version, p.requires_grad) for p in params) + (capturing,)
    hit = _FLAT_CACHE.get(owner)          # kept outside the module: a graph tensor in its __dict__ would break deepcopy
    if hit is not None and hit[0] == key and not hit[2]["consumed"]:
        st = hit[2]
        ok = True
        if flat_is_cuda(hit[1]) and torch.cuda.current_stream() != st["stream"]:
            try:
                torch.cuda.current_stream().wait_event(st["event"])   # made on another stream: order this one after it
            except RuntimeError:          # e.g. an event left behind by an aborted capture: rebuild
                ok = False
        if ok:
            return hit[1]
    flat = torch.cat([p.reshape(-1).float() for p in params])
    state = {"consumed": False}
    if flat_is_cuda(flat):
        state["stream"] = torch.cuda.current_stream()
        state["event"] = torch.cuda.Event()
        state["event"].record(state["stream"])
    if flat.requires_grad:
        def _mark(_g, state=state):
            state["consumed"] = True
        flat.register_hook(_mark)
    _FLAT_CACHE[owner] = (key, flat, state)
    return flat


def _xchg_args(xchg):
    if xchg is None or xchg.world <= 1:
        return None, 0, 1
    return ctypes.addressof(xchg.peers), xchg.rank, xchg.world


class ComplexityTrainFn(torch.autograd.Function):
    """phi (B,ht,wt,8) -> complexity (B,ht,wt): MLP + 5x5 bilateral + clamp (morphology.py:959-968)."""

    @staticmethod
    def forward(ctx, phi, flat, eval_block, consts):
        out, raw = ops.complexity(phi, eval_block, consts, want_raw=True)
        ctx.save_for_backward(phi, raw, flat)
        return out

    @staticmethod
    def backward(ctx, g):
        phi, raw, flat = ctx.saved_tensors
        B, ht, wt, _ = phi.shape
        N = B * ht * wt
        lib = _lib.load()
        scratch = torch.empty((int(lib.mcaq_cmlp_train_scratch_floats(N)),), device=phi.device, dtype=torch.float32)
        gcraw = torch.empty((N,), device=phi.device, dtype=torch.float32)
        gP = torch.zeros((CMLP_PARAMS,), device=phi.device, dtype=torch.float32)
        g = g.contiguous().float()
        ops._call("mcaq_complexity_train_bwd", phi.data_ptr(), raw.data_ptr(), g.data_ptr(), B, ht, wt, flat.data_ptr(),
                  scratch.data_ptr(), gcraw.data_ptr(), gP.data_ptr(), ops._stream())
        return None, gP, None, None


class MapperTrainFn(torch.autograd.Function):
    """complexity (N rows) -> continuous bits with TRAIN-MODE BatchNorm1d (bit_allocation.py:218-280): batch statistics
    over all rows (all ranks when `xchg` is a peer exchange of world > 1), running statistics updated in place."""

    @staticmethod
    def forward(ctx, c, flat, bns, temperature, lo, hi, xchg):
        c = c.contiguous().float()
        N = c.numel()
        lib = _lib.load()
        dev = c.device
        scratch = torch.empty((int(lib.mcaq_mapper_train_scratch_floats(N)),), device=dev, dtype=torch.float32)
        stats = torch.empty((256,), device=dev, dtype=torch.float32)
        bits = torch.empty_like(c)
        use_t = temperature is not None
        t = max(float(temperature), 0.1) if use_t else 1.0
        run, nbt = [], []
        for bn in bns:
            track = bn.track_running_stats and bn.running_mean is not None
            run += [bn.running_mean.data_ptr() if track else None, bn.running_var.data_ptr() if track else None]
            nbt.append(bn.num_batches_tracked.data_ptr() if track and bn.num_batches_tracked is not None else None)
        mom = bns[0].momentum if bns[0].momentum is not None else 0.1
        peers, rank, world = _xchg_args(xchg)
        ops._call("mcaq_mapper_train_fwd", c.data_ptr(), N, flat.data_ptr(), t, int(use_t), float(lo), float(hi),
                  scratch.data_ptr(), stats.data_ptr(), *run, *nbt, float(mom), float(bns[0].eps), bits.data_ptr(),
                  peers, rank, world, ops._stream())          # num_batches_tracked += 1 happens in the kernel
        ctx.save_for_backward(c, flat, scratch, stats)
        ctx.cfg = (t, int(use_t), float(lo), float(hi), float(bns[0].eps), xchg)
        return bits

    @staticmethod
    def backward(ctx, g):
        c, flat, scratch, stats = ctx.saved_tensors
        t, use_t, lo, hi, eps, xchg = ctx.cfg
        N = c.numel()
        gc = torch.empty_like(c)
        gP = torch.zeros((MAPPER_PARAMS,), device=c.device, dtype=torch.float32)
        g = g.contiguous().float()
        peers, rank, world = _xchg_args(xchg)
        ops._call("mcaq_mapper_train_bwd", c.data_ptr(), N, flat.data_ptr(), t, use_t, lo, hi, scratch.data_ptr(),
                  stats.data_ptr(), eps, g.data_ptr(), gc.data_ptr(), gP.data_ptr(), peers, rank, world, ops._stream())
        return gc, gP, None, None, None, None, None


class SoftMaskTrainFn(torch.autograd.Function):
    """(bit_map (B,Ht,Wt), sum_c|x| plane) -> m (B,H,W) (quantization.py:213-239); gradients to the bit map and the net."""

    @staticmethod
    def forward(ctx, bit_map, flat, abs_plane, C, packed):
        bit_map = bit_map.contiguous().float()
        B, H, W = abs_plane.shape
        Ht, Wt = int(bit_map.shape[1]), int(bit_map.shape[2])
        m = ops.soft_mask(bit_map, abs_plane, C, packed)
        act = torch.empty((B, Ht, Wt), device=abs_plane.device, dtype=torch.float32)
        ops._call("mcaq_softmask_act", abs_plane.data_ptr(), B, int(C), H, W, Ht, Wt, act.data_ptr(), ops._stream())
        ctx.save_for_backward(bit_map, act, packed)
        ctx.geom = (B, H, W, Ht, Wt)
        return m

    @staticmethod
    def backward(ctx, gm):
        bit_map, act, packed = ctx.saved_tensors
        B, H, W, Ht, Wt = ctx.geom
        gm = gm.contiguous().float()
        dbit = torch.empty_like(bit_map)
        gP = torch.zeros((SOFTMASK_PARAMS,), device=bit_map.device, dtype=torch.float32)
        ops._call("mcaq_softmask_train_bwd", gm.data_ptr(), bit_map.data_ptr(), act.data_ptr(), packed.data_ptr(), B, H, W,
                  Ht, Wt, dbit.data_ptr(), gP.data_ptr(), ops._stream())
        return dbit, gP, None, None, None


class BitStatsFn(torch.autograd.Function):
    """bit_map (B,ht,wt) -> tensor [sum(b), TV(b)] (fp32, 2 elements): the reductions behind avg_bits / Lbit and
    Lsmooth (models/mcaq_yolo.py:86-118, 575) in one launch each way."""

    @staticmethod
    def forward(ctx, bit_map):
        bm = bit_map.contiguous().float()
        B, ht, wt = bm.shape
        out = torch.zeros((2,), device=bm.device, dtype=torch.float32)
        ops._call("mcaq_bit_stats", bm.data_ptr(), B, ht, wt, out.data_ptr(), None, None, ops._stream())
        ctx.save_for_backward(bm)
        return out

    @staticmethod
    def backward(ctx, g):
        bm, = ctx.saved_tensors
        B, ht, wt = bm.shape
        grad = torch.empty_like(bm)
        w = g.contiguous().float()
        ops._call("mcaq_bit_stats", bm.data_ptr(), B, ht, wt, None, w.data_ptr(), grad.data_ptr(), ops._stream())
        return grad


class BitLossesFn(torch.autograd.Function):
    """(bit maps of S <= 4 scales) -> tensor [avg_bits, Lbit, Lsmooth] in ONE launch each way (mcaq_bit_losses): the
    single-rank form of `bit_map_losses`."""

    @staticmethod
    def _args(maps):
        S = len(maps)
        ptrs = (ctypes.c_void_p * S)(*[m.data_ptr() for m in maps])
        Bs = (ctypes.c_int * S)(*[m.shape[0] for m in maps])
        hs = (ctypes.c_int * S)(*[m.shape[1] for m in maps])
        ws = (ctypes.c_int * S)(*[m.shape[2] for m in maps])
        return S, ptrs, Bs, hs, ws

    @staticmethod
    def forward(ctx, target, *bit_maps):
        maps = [b.contiguous().float() for b in bit_maps]
        S, ptrs, Bs, hs, ws = BitLossesFn._args(maps)
        out = torch.empty((3 + 2 * S,), device=maps[0].device, dtype=torch.float32)
        ops._call("mcaq_bit_losses", ptrs, Bs, hs, ws, S, float(target), out.data_ptr(), None, None, ops._stream())
        ctx.save_for_backward(out, *maps)
        ctx.target = float(target)
        return out[:3]

    @staticmethod
    def backward(ctx, g):
        out, *maps = ctx.saved_tensors
        S, ptrs, Bs, hs, ws = BitLossesFn._args(maps)
        grads = [torch.empty_like(m) for m in maps]
        gptrs = (ctypes.c_void_p * S)(*[t.data_ptr() for t in grads])
        g = g.contiguous().float()
        ops._call("mcaq_bit_losses", ptrs, Bs, hs, ws, S, ctx.target, out.data_ptr(), g.data_ptr(), gptrs, ops._stream())
        return (None, *grads)


def bit_map_losses(bit_maps, target_bits: float, group=None):
    """(avg_bits, Lbit, Lsmooth) of the per-scale bit maps exactly as MCAQYOLO.forward / MCAQYOLOLoss form them
    (models/mcaq_yolo.py:575, 86-118): avg_bits = mean over scales of the per-scale mean, Lsmooth = mean over scales
    of the per-edge mean total variation.  With a process group the per-scale means are taken over the GLOBAL batch
    (one all-reduce of 2 x len(bit_maps) floats; the gradient stays local and is scaled by the global count)."""
    import torch.distributed as dist
    sharded = group is not None or (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1)
    if not sharded and 1 <= len(bit_maps) <= 4 and all(b.is_cuda and b.dim() == 3 for b in bit_maps):
        r = BitLossesFn.apply(float(target_bits), *bit_maps)                 # one launch forward, one backward
        return r[0], r[1], r[2]
    stats = torch.stack([BitStatsFn.apply(b) for b in bit_maps])             # (S, 2)
    ckey = (tuple(tuple(b.shape) for b in bit_maps), stats.device)
    counts = _COUNTS.get(ckey)          # cached: no host-to-device copy in the steady state (graph-capturable)
    if counts is None:
        counts = torch.tensor([[b.numel(), b.shape[0] * ((b.shape[1] - 1) * b.shape[2] + b.shape[1] * (b.shape[2] - 1))]
                               for b in bit_maps], device=stats.device, dtype=torch.float32)
        _COUNTS[ckey] = counts
    import torch.distributed as dist
    if group is not None or (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        stats = _AllReduceSum.apply(stats, group)
        counts = counts * dist.get_world_size(group)
    means = stats / counts.clamp(min=1.0)
    avg_bits = means[:, 0].mean()
    return avg_bits, (avg_bits - target_bits) ** 2, means[:, 1].mean()


class _AllReduceSum(torch.autograd.Function):
    """SUM all-reduce whose backward is the identity on the local contribution ("straight-through local
    gradient"): every rank differentiates the same global scalar w.r.t. its own rows."""

    @staticmethod
    def forward(ctx, x, group):
        import torch.distributed as dist
        y = x.clone()
        dist.all_reduce(y, op=dist.ReduceOp.SUM, group=group)
        return y

    @staticmethod
    def backward(ctx, g):
        return g, None


def allreduce_grads(modules, group=None, average: bool = False):
    """One flat all-reduce (SUM, or mean with `average`) of the gradients of the hot path's small networks --
    complexity MLP (2881), mapper (4609), soft-mask nets (170 each): ~32 KB -- instead of one per tensor."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
        return
    ps = [p for m in modules for p in m.parameters() if p.grad is not None]
    if not ps:
        return
    flat = torch.cat([p.grad.reshape(-1).float() for p in ps])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for p in ps:
        n = p.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n


def merge_batch_stats(parts):
    """Rank-ordered Chan merge of per-rank (count, mean, M2) -- the host restatement of what the mapper kernels do
    over peer memory (csrc/train_nets.cu: batch_stats); used by the gloo test."""
    cnt, mean, m2 = 0.0, None, None
    for n, mu, q in parts:
        if n <= 0:
            continue
        if mean is None:
            cnt, mean, m2 = float(n), mu.clone(), q.clone()
            continue
        tot = cnt + n
        d = mu - mean
        mean = mean + d * (n / tot)
        m2 = m2 + q + d * d * (cnt * n / tot)
        cnt = tot
    return cnt, mean, m2
