"""Host-side mirror of the reference's quantizer interface, on the sm_100a kernels.

Drop-in for ``som_quantizer.ResidualQuantizer`` (third-party quantization-maps package imported at
``/root/reference/networks/vae.py:6``): same constructor keywords (``vae.py:245-251``), same call and
return order ``(x_quantized, index, commit_loss)`` (``vae.py:315-318``), same attributes the reference
reads (``.num_quantizers`` ``training.py:183``, ``.use_som`` ``utils.py:239``, ``.quantizers[i].dequantize``
``vae.py:333``, ``.quantizers[i].som.height/.width`` ``utils.py:244-245``, ``.get_stale_clusters()``
``training.py:435``, ``.update_cutoff()`` ``vae.py:350-351``).

PyTorch owns every tensor (device memory, streams, ``torch.distributed``); all arithmetic of the path
runs in ``librvq_sm100a.so`` through the C ABI of ``include/rvq_sm100a.h``.  There is no CPU path:
a CPU tensor or a missing library raises.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List, Optional

import torch
from torch import nn

from . import _lib
from ._lib import (RVQError, RVQ_ALGO_EXACT_SCAN, RVQ_ALGO_TENSOR, RVQ_FLAG_CLUSTER_SHIFT, RVQ_FLAG_COUNTERS,
                   RVQ_KERNELS)

EMA_DECAY = 0.99   # ASSUMED (rosinality / Jukebox lineage, README.md:26-27)
EMA_EPS = 1e-5     # ASSUMED
SOM_SHRINK = 0.1   # ASSUMED: SOM neighbourhood width 1 / (1 + SOM_SHRINK * update steps) (arXiv 2302.07950, README.md:10)
SOM_MAX_RADIUS = 4
_M64 = (1 << 64) - 1


def tuple_checker(item, length):
    """Same behaviour as ``/root/reference/networks/utils.py:212-220``."""
    if isinstance(item, (int, float, str)):
        item = [item] * length
    elif isinstance(item, (tuple, list)):
        assert len(item) == length, f"Expected tuple of length {length}, got {len(item)}"
    return item


def approximate_square_root(n: int):
    h = int(math.isqrt(int(n)))
    while h > 1 and n % h:
        h -= 1
    return h, n // h


def som_weights(kernel_type: str, t: int, shrink: float = SOM_SHRINK):
    """(radius, row-major weights) of the SOM neighbourhood after ``t`` updates: "hard" = the code itself 1 and its
    four grid neighbours sigma_t = 1 / (1 + shrink t); "gaussian" = exp(-dist^2 / (2 sigma_t^2)) within
    radius min(4, max(1, ceil(3 sigma_t)))."""
    sigma = 1.0 / (1.0 + float(shrink) * float(t))
    if kernel_type == "hard":
        return 1, [0.0, sigma, 0.0, sigma, 1.0, sigma, 0.0, sigma, 0.0]
    if kernel_type == "gaussian":
        r = min(SOM_MAX_RADIUS, max(1, int(math.ceil(3.0 * sigma))))
        return r, [math.exp(-(dy * dy + dx * dx) / (2.0 * sigma * sigma))
                   for dy in range(-r, r + 1) for dx in range(-r, r + 1)]
    raise ValueError(f"som_kernel_type must be 'hard' or 'gaussian', got {kernel_type!r}")


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _frame_addressing(x: torch.Tensor):
    """(x3, N, L, stride_b, stride_l, stride_d) for a (..., L, d) tensor; copies only when the layout
    cannot be addressed as base + b*sb + l*sl + i*sd with a dense, non-overlapping footprint."""
    if x.dim() == 2:
        x3 = x.unsqueeze(0)
    elif x.dim() == 3:
        x3 = x
    else:
        x3 = x.reshape(-1, x.shape[-2], x.shape[-1])
    ok = x3.is_contiguous() or x3.transpose(1, 2).is_contiguous()
    if ok and x3.stride(2) == 1:
        ok = (x3.stride(1) % 4 == 0) and (x3.stride(0) % 4 == 0) and (x3.data_ptr() % 16 == 0)
    if not ok:
        # .contiguous() is a no-op on a tensor that already is contiguous but sits at an unaligned offset (a view into
        # a larger buffer) or whose strides of size-1 dimensions are arbitrary: force a fresh, aligned copy
        x3 = x3.clone(memory_format=torch.contiguous_format)
    B, L, _ = x3.shape
    return x3, B * L, L, x3.stride(0), x3.stride(1), x3.stride(2)


class _SOMGrid:
    """Grid shape of the self-organising map attached to a stage (``utils.py:244-245,257``)."""

    def __init__(self, K: int):
        self.height, self.width = approximate_square_root(K)


class _Stage:
    """``ResidualQuantizer.quantizers[i]``: one stage's view of the shared state."""

    def __init__(self, parent: "ResidualQuantizer", q: int):
        self._p, self._q = parent, q
        if parent.use_som:
            self.som = _SOMGrid(parent.codebook_sizes[q])

    @property
    def codebook(self) -> torch.Tensor:
        return self._p.codebooks[self._q, : self._p.codebook_sizes[self._q]]

    def dequantize(self, idx: torch.Tensor) -> torch.Tensor:
        """Code lookup ``(..., ) int64 -> (..., d)`` (``vae.py:333``)."""
        return self._p.dequantize(idx.unsqueeze(-1), first_stage=self._q)


class _RVQFunction(torch.autograd.Function):
    """Straight-through quantization + commit loss with hand-written backward.

    forward:  xq (kernel), out = x + (xq - x)            [value of x + (xq - x).detach()]
              commit = w * sum_q mean((r_q - sg z_q)^2) (+ sum_q mean((sg r_q - z_q)^2) for "base")
    backward: d out / d x = I ; d commit / d x = w * 2/(N d) * sum_q r_{q+1} ;
              d commit / d C_q[k] = -2/(N d) * sum_{n: idx=k} r_{q+1}[n]   ("base" only)
    """

    @staticmethod
    def forward(ctx, x, codebooks, mod, nq, update):
        # With update=True the codebook maintenance (EMA refresh, stale-code re-seeding) rewrites `codebooks` through
        # raw pointers right after the encode: autograd cannot see that write, and backward re-walks the residual
        # chain r_{q+1} = r_q - C_q[idx].  So the codebooks the indices were computed with are snapshotted first.
        # SOM width of THIS call (the update below increments the step counter)
        som_t = mod._update_step_index() if (mod.use_som and mod.quantizer_class == "base") else None
        xq, idx, commit_sq, cb_used = mod._encode(x, nq, update, snapshot=update)
        N = idx.numel() // nq
        w = mod.commitment_weight + (1.0 if mod.quantizer_class == "base" else 0.0)
        commit = (commit_sq.sum() * (w / max(N * mod.dim, 1))).to(torch.float32)
        ctx.mod, ctx.nq = mod, nq
        ctx.som_t = som_t
        ctx.save_for_backward(x, idx, codebooks if cb_used is None else cb_used)
        ctx.mark_non_differentiable(idx)
        out = x + (xq - x)
        return out, idx, commit

    @staticmethod
    def backward(ctx, g_out, _g_idx, g_commit):
        x, idx, codebooks = ctx.saved_tensors
        mod, nq = ctx.mod, ctx.nq
        need_x = ctx.needs_input_grad[0]
        need_cb = ctx.needs_input_grad[1] and mod.quantizer_class == "base"
        if g_commit is None or not (need_x or need_cb):
            return (g_out if need_x else None), None, None, None, None
        gx, gcb = mod._backward(x, idx, codebooks.detach(), nq, g_out if need_x else None, g_commit, need_x, need_cb,
                                som_t=ctx.som_t)
        return gx, gcb, None, None, None


def _forget_step_mirror(module, _incompatible_keys):
    """load_state_dict post-hook: the host mirror of ``update_steps`` is re-read from the loaded buffer."""
    module._steps_host = None


class ResidualQuantizer(nn.Module):
    """Residual vector quantizer on B200 (see module docstring for the interface it mirrors)."""

    def __init__(self, num_quantizers, dim, quantizer_class="ema", codebook_sizes=1024,
                 vq_cutoff_freq=1, use_som=True, som_kernel_type="hard",
                 decay=EMA_DECAY, eps=EMA_EPS, commitment_weight=1.0, algo="tensor",
                 som_shrink=SOM_SHRINK, reseed_seed=0, kernel="auto", cluster=0):
        super().__init__()
        if quantizer_class not in ("ema", "base"):
            raise ValueError(f"quantizer_class must be 'ema' or 'base', got {quantizer_class!r}")
        self.num_quantizers = int(num_quantizers)
        self.dim = int(dim)
        self.quantizer_class = quantizer_class
        self.codebook_sizes: List[int] = [int(k) for k in tuple_checker(codebook_sizes, self.num_quantizers)]
        self.vq_cutoff_freq = float(vq_cutoff_freq)
        self.use_som = bool(use_som)
        self.som_kernel_type = som_kernel_type
        if self.use_som:
            som_weights(som_kernel_type, 0)        # validates the kernel type
        self.som_shrink, self.reseed_seed = float(som_shrink), int(reseed_seed)
        self.decay, self.eps, self.commitment_weight = float(decay), float(eps), float(commitment_weight)
        self.algo = algo
        # launch options of the fused kernel (tests cross-check the kernels against each other; "auto" is the product)
        self.kernel, self.cluster, self.counters = kernel, int(cluster), False
        K = max(self.codebook_sizes)
        self.K = K
        cb = torch.randn(self.num_quantizers, K, self.dim)   # ASSUMED init
        if quantizer_class == "base":
            self.codebooks = nn.Parameter(cb)
        else:
            self.register_buffer("codebooks", cb)
        self.register_buffer("ema_count", torch.ones(self.num_quantizers, K))
        self.register_buffer("ema_sum", cb.detach().clone())
        self.register_buffer("k_valid", torch.tensor(self.codebook_sizes, dtype=torch.int32))
        self.register_buffer("update_steps", torch.zeros((), dtype=torch.int64))   # EMA updates applied so far
        self.register_buffer("n_replaced", torch.zeros(self.num_quantizers, dtype=torch.int32),
                             persistent=False)                                     # codes re-seeded by the last update
        self._steps_host = None    # host mirror of update_steps (read back once, then counted on the host)
        self.register_load_state_dict_post_hook(_forget_step_mirror)
        self.quantizers = [_Stage(self, q) for q in range(self.num_quantizers)]
        self._derived = None       # (key, cb_op, cb_norm, cb_meta)
        self._derived_bufs = None  # the three buffers, reused across rebuilds (stream-ordered: one stream per module)
        self._ws = None
        self._stats = None
        self._spread = None
        self.last_stats = None     # flat [sum | cnt | replacement vectors] of the most recent update (for inspection)
        self.kernel_events = None  # a list collects (start, stop) CUDA events around every rvq_encode launch (bench.py)
        self.comm_events = None    # ... around every all-reduce of the statistics
        self.update_events = None  # ... and around the whole codebook maintenance that follows the kernel
        self.comm_copy = False     # True: all-reduce a copy of the statistics instead of the kernel's own buffer
        self._comm = None
        self.sync_stats = True     # False: every rank updates from its own shard only (replicas DIVERGE; measurements)

    # ------------------------------------------------------------------ derived operands / scratch
    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._derived = self._derived_bufs = self._ws = self._stats = self._spread = self._steps_host = None
        return out

    def _check_device(self, t: torch.Tensor):
        if not t.is_cuda:
            raise RVQError("ResidualQuantizer runs on sm_100a only: got a CPU tensor (there is no CPU path)")
        if self.codebooks.device != t.device:
            raise RVQError(f"input on {t.device} but codebooks on {self.codebooks.device}")

    def invalidate(self):
        """Call after writing ``codebooks`` in place outside this class by means that do not bump the tensor's version
        counter while the module is in eval mode (nothing in the reference does)."""
        self._derived = None

    def train(self, mode: bool = True):
        # whatever a fused optimizer wrote while training is picked up by the first eval-mode call (training.py:494-499)
        if mode != self.training:
            self._derived = None
        return super().train(mode)

    def _prepared(self):
        """fp16 operands / norms / metadata of the current codebooks (K0), rebuilt when the codebooks changed.

        Buffers ("ema") only change through this class.  Parameters ("base") are rewritten in place by the optimizer:
        foreach / for-loop optimizers bump the tensor's version counter, but FUSED ones (torch._fused_adam_) do not
        (measured: tests/test_gpu_hardening.py), so in training mode gradient-trained codebooks are re-prepared on every
        call (two small HBM-bound kernels over nq K d floats); in eval mode the (pointer, version) key is used."""
        cb = self.codebooks.detach()
        key = (cb.data_ptr(), cb._version, str(cb.device))
        always = self.quantizer_class == "base" and self.training
        if self._derived is None or self._derived[0] != key or always:
            lib = _lib.load()
            nq, K, d = self.num_quantizers, self.K, self.dim
            bufs = self._derived_bufs
            if bufs is None or bufs[0].device != cb.device:
                ob, nb, mb = C.c_size_t(), C.c_size_t(), C.c_size_t()
                _lib.check(lib.rvq_prepared_bytes(nq, K, d, C.byref(ob), C.byref(nb), C.byref(mb)), "rvq_prepared_bytes")
                bufs = (torch.empty(ob.value // 2, dtype=torch.float16, device=cb.device),
                        torch.empty(nb.value // 4, dtype=torch.float32, device=cb.device),
                        torch.empty(mb.value // 4, dtype=torch.float32, device=cb.device))
                self._derived_bufs = bufs
            op, nrm, meta = bufs
            with torch.cuda.device(cb.device):
                _lib.check(lib.rvq_prepare_codebooks(_ptr(cb), _ptr(self.k_valid), nq, K, d, _ptr(op), _ptr(nrm),
                                                     _ptr(meta), _stream()), "rvq_prepare_codebooks")
            self._derived = (key, op, nrm, meta)
        return self._derived[1:]

    def _workspace(self, device):
        if self._ws is None or self._ws.device != device:
            lib = _lib.load()
            n = C.c_size_t()
            _lib.check(lib.rvq_workspace_bytes(self.num_quantizers, self.K, self.dim, 0, C.byref(n)),
                       "rvq_workspace_bytes")
            self._ws = torch.empty(n.value, dtype=torch.uint8, device=device)
        return self._ws

    def _stats_buffers(self, device):
        """One flat fp32 buffer [nq K d sums | nq K counts | nq K d replacement vectors] = one all-reduce payload."""
        nq, K, d = self.num_quantizers, self.K, self.dim
        if self._stats is None or self._stats.device != device:
            # zeros once: the replacement-vector part of stages a partial call (n < num_quantizers) does not write
            # travels through the all-reduce as zeros, not as uninitialised memory
            self._stats = torch.zeros(2 * nq * K * d + nq * K, dtype=torch.float32, device=device)
            self._spread = torch.empty(nq * K * d + nq * K, dtype=torch.float32, device=device)
        flat = self._stats
        return flat, flat[: nq * K * d], flat[nq * K * d: nq * K * (d + 1)], flat[nq * K * (d + 1):]

    def _update_step_index(self) -> int:
        if self._steps_host is None:
            self._steps_host = int(self.update_steps)
        return self._steps_host

    def _update_codebooks(self, x3, N, L, sb, sl, sd, nq, idx, flat, ssum, scnt, rep):
        """All-reduce of the statistics -> SOM neighbourhood -> EMA refresh -> stale-code re-seeding (K3 and
        SURVEY 8f rows 2, 3), on the current stream.

        quantizer_class "base" (the reference's config/training.yml:21 passes update_codebook=True for it too,
        training.py:305-308): the codebooks are trained by gradient, so only the usage counts are averaged
        (get_stale_clusters / update_cutoff stay meaningful, training.py:435,454,461) and stale codes are re-seeded;
        the SOM neighbourhood acts on the codebook gradient instead (``_backward``).  ASSUMED semantics."""
        lib = _lib.load()
        K, d = self.K, self.dim
        ema = self.quantizer_class == "ema"
        cb = self.codebooks.detach()
        dist = torch.distributed
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() and self.sync_stats else 1
        rank = dist.get_rank() if world > 1 else 0
        t = self._update_step_index()
        cutoff = self.vq_cutoff_freq
        reseed = cutoff > 0
        nsum, ncnt = self.num_quantizers * K * d, self.num_quantizers * K
        if reseed:
            # replacement vectors need the codebooks `idx` was computed with: gathered BEFORE the EMA refresh; frames
            # are numbered globally (equal shards per rank), non-owners contribute zeros to the sum
            seed = (self.reseed_seed + t * 0xD1B54A32D192ED03) & _M64
            _lib.check(lib.rvq_reseed_gather(_ptr(x3), N, L, sb, sl, sd, d, nq, K, _ptr(cb), _ptr(idx),
                                             _ptr(self.ema_count), self.decay, cutoff, seed, rank * N, N * world,
                                             _ptr(rep), _stream()), "rvq_reseed_gather")
        # one flat payload [sum | cnt | rep]: "ema" sends [sum | cnt (| rep)], "base" only [cnt (| rep)]
        lo = 0 if ema else nsum
        hi = nsum + ncnt + (nsum if reseed else 0)
        if world > 1:
            # the only place the path crosses frame shards: one SUM all-reduce
            if self.comm_events is not None:
                e0 = torch.cuda.Event(enable_timing=True)
                e0.record()
            if self.comm_copy:
                # all-reduce a COPY: the buffer the encode kernel reduces into never becomes NCCL's send/receive buffer
                if self._comm is None or self._comm.numel() != flat.numel() or self._comm.device != flat.device:
                    self._comm = torch.empty_like(flat)
                self._comm[lo:hi].copy_(flat[lo:hi])
                dist.all_reduce(self._comm[lo:hi], op=dist.ReduceOp.SUM)
                flat = self._comm
                nsd = self.num_quantizers * K * d
                ssum, scnt, rep = flat[:nsd], flat[nsd: nsd + self.num_quantizers * K], flat[nsd + self.num_quantizers * K:]
            else:
                dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM)
            if self.comm_events is not None:
                e1 = torch.cuda.Event(enable_timing=True)
                e1.record()
                self.comm_events.append((e0, e1))
        if ema:
            if self.use_som:
                radius, w = som_weights(self.som_kernel_type, t, self.som_shrink)
                hw = []
                for q in range(nq):
                    hw += list(approximate_square_root(self.codebook_sizes[q]))
                ssum2, scnt2 = self._spread[:nsum], self._spread[nsum:]
                _lib.check(lib.rvq_som_spread(_ptr(ssum), _ptr(scnt), _ptr(ssum2), _ptr(scnt2), (C.c_int * len(hw))(*hw),
                                              nq, K, d, radius, (C.c_float * len(w))(*w), _stream()), "rvq_som_spread")
                ssum, scnt = ssum2, scnt2
            _lib.check(lib.rvq_ema_finalize(_ptr(cb), _ptr(self.ema_count), _ptr(self.ema_sum), _ptr(ssum), _ptr(scnt),
                                            _ptr(self.k_valid), nq, K, d, self.decay, self.eps, _stream()),
                       "rvq_ema_finalize")
        else:
            _lib.check(lib.rvq_ema_counts(_ptr(self.ema_count), _ptr(scnt), _ptr(self.k_valid), nq, K, self.decay,
                                          _stream()), "rvq_ema_counts")
        if reseed:
            _lib.check(lib.rvq_reseed_apply(_ptr(cb), _ptr(self.ema_count), _ptr(self.ema_sum), _ptr(rep),
                                            _ptr(self.k_valid), nq, K, d, cutoff, cutoff, _ptr(self.n_replaced),
                                            _stream()), "rvq_reseed_apply")
        self.update_steps += 1
        self._steps_host = t + 1
        if ema or reseed:
            self._derived = None     # codebooks changed: operands are rebuilt before the next call
        self.last_stats = flat

    # ------------------------------------------------------------------ the hot path
    def _encode(self, x: torch.Tensor, nq: int, update: bool, ws: Optional[torch.Tensor] = None,
                snapshot: bool = False):
        """Run K1(+K2) [+ all-reduce + K3] on ``x`` (..., L, d); returns (xq like x, idx (..., L, nq), commit_sq,
        copy of the codebooks the indices refer to if ``snapshot`` and the call updated them, else None)."""
        lib = _lib.load()
        self._check_device(x)
        if x.dtype != torch.float32:
            x = x.float()
        x = x.detach()
        x3, N, L, sb, sl, sd = _frame_addressing(x)
        op, nrm, meta = self._prepared()
        cb = self.codebooks.detach()
        dev = x3.device
        xq = torch.empty_strided(x3.shape, x3.stride(), dtype=torch.float32, device=dev)
        idx = torch.empty((N, nq), dtype=torch.int64, device=dev)
        commit_sq = torch.empty(nq, dtype=torch.float64, device=dev)
        if ws is None:
            ws = self._workspace(dev)
        ssum = scnt = flat = rep = None
        if update:
            flat, ssum, scnt, rep = self._stats_buffers(dev)
            flat[: ssum.numel() + scnt.numel()].zero_()
        flags = RVQ_ALGO_EXACT_SCAN if self.algo == "exact_scan" else RVQ_ALGO_TENSOR
        flags |= RVQ_KERNELS[self.kernel] | (self.cluster << RVQ_FLAG_CLUSTER_SHIFT)
        if self.counters:
            flags |= RVQ_FLAG_COUNTERS
        with torch.cuda.device(dev):
            if self.kernel_events is not None:
                k0 = torch.cuda.Event(enable_timing=True)
                k0.record()
            _lib.check(lib.rvq_encode(_ptr(x3), N, L, sb, sl, sd, self.dim, nq, self.K, _ptr(cb), _ptr(op), _ptr(nrm),
                                      _ptr(meta), _ptr(xq), _ptr(idx), _ptr(commit_sq), _ptr(ssum), _ptr(scnt),
                                      _ptr(ws), ws.numel(), flags, _stream()), "rvq_encode")
            if self.kernel_events is not None:
                k1 = torch.cuda.Event(enable_timing=True)
                k1.record()
                self.kernel_events.append((k0, k1))
            cb_used = None
            if update:
                if snapshot:
                    cb_used = cb.clone()
                self._update_step_index()
                self._update_codebooks(x3, N, L, sb, sl, sd, nq, idx, flat, ssum, scnt, rep)
                if self.update_events is not None and self.kernel_events is not None:
                    k2 = torch.cuda.Event(enable_timing=True)
                    k2.record()
                    self.update_events.append((k1, k2))
        xq = xq.reshape(x.shape) if xq.shape != x.shape else xq
        return xq, idx.reshape(*x.shape[:-1], nq), commit_sq, cb_used

    def read_counters(self) -> List[int]:
        """The 32 event counters of the last launch made with ``self.counters = True`` (last 256 bytes of the
        workspace; synchronises)."""
        ws = self._ws
        if ws is None:
            return [0] * 32
        off = (ws.numel() - 256) & ~7
        return ws[off: off + 256].view(torch.int64).cpu().tolist()

    def _backward(self, x, idx, cb, nq, g_out, g_commit, need_x, need_cb, som_t=None):
        """rvq_backward: gx = g_out + g_commit * w * 2/(N d) * sum_q r_{q+1}; gcb[q, idx] -= g_commit * 2/(N d) * r_{q+1};
        with ``use_som`` the codebook gradient is spread over each code's map neighbourhood (rvq_som_spread), so the
        neighbours of a winner follow it - the SOM coupling for gradient-trained ("base") codebooks, ASSUMED."""
        lib = _lib.load()
        xf = x.detach()
        if xf.dtype != torch.float32:
            xf = xf.float()
        x3, N, L, sb, sl, sd = _frame_addressing(xf)
        dev = x3.device
        gx = torch.empty_strided(x3.shape, x3.stride(), dtype=torch.float32, device=dev) if need_x else None
        go = add_later = None
        if need_x and g_out is not None:
            g3 = g_out.reshape(x3.shape)
            if g3.dtype == torch.float32 and g3.stride() == x3.stride() and g3.data_ptr() % 16 == 0:
                go = g3                      # same addressing as x: fused into the kernel
            else:
                add_later = g3
        gcb = torch.zeros_like(cb) if need_cb else None
        gc = g_commit.detach().to(device=dev, dtype=torch.float32).contiguous()
        with torch.cuda.device(dev):
            _lib.check(lib.rvq_backward(_ptr(x3), N, L, sb, sl, sd, self.dim, nq, self.K, _ptr(cb), _ptr(idx.reshape(-1, nq)),
                                        _ptr(go), _ptr(gc), self.commitment_weight, 1.0 if need_cb else 0.0,
                                        _ptr(gx), _ptr(gcb), _stream()), "rvq_backward")
            if gcb is not None and self.use_som:
                t = self._update_step_index() if som_t is None else som_t
                radius, w = som_weights(self.som_kernel_type, t, self.som_shrink)
                hw = []
                for q in range(nq):
                    hw += list(approximate_square_root(self.codebook_sizes[q]))
                g2 = torch.zeros_like(gcb)
                zc = torch.zeros(2, self.num_quantizers, self.K, dtype=torch.float32, device=dev)
                _lib.check(lib.rvq_som_spread(_ptr(gcb), _ptr(zc[0]), _ptr(g2), _ptr(zc[1]), (C.c_int * len(hw))(*hw),
                                              nq, self.K, self.dim, radius, (C.c_float * len(w))(*w), _stream()),
                           "rvq_som_spread")
                gcb = g2
        if gx is not None:
            if add_later is not None:
                gx = gx + add_later
            gx = gx.reshape(x.shape).to(x.dtype)
        return gx, gcb

    def forward(self, x, n=None, update_codebook=False, prioritize_early=False):
        if prioritize_early:
            raise NotImplementedError("prioritize_early=True: semantics unknown (never used by the reference)")
        nq = self.num_quantizers if n is None else int(n)
        if not 1 <= nq <= self.num_quantizers:
            raise ValueError(f"n must be in [1, {self.num_quantizers}], got {nq}")
        if x.shape[-1] != self.dim:
            raise ValueError(f"last dimension must be {self.dim}, got {tuple(x.shape)}")
        update = bool(update_codebook) and self.training     # "base" keeps usage counts / re-seeds too (A13)
        if torch.is_grad_enabled() and (x.requires_grad or (self.quantizer_class == "base" and
                                                            self.codebooks.requires_grad)):
            out, idx, commit = _RVQFunction.apply(x, self.codebooks, self, nq, update)
            return out, idx, commit
        xq, idx, commit_sq, _ = self._encode(x, nq, update)
        N = idx.numel() // nq
        w = self.commitment_weight + (1.0 if self.quantizer_class == "base" else 0.0)
        commit = (commit_sq.sum() * (w / max(N * self.dim, 1))).to(torch.float32)
        return xq, idx, commit

    # ------------------------------------------------------------------ decode side
    def dequantize(self, idx: torch.Tensor, first_stage: int = 0) -> torch.Tensor:
        """Sum of code vectors: idx (..., n) int64 -> (..., d) using stages first_stage .. first_stage+n-1."""
        lib = _lib.load()
        self._check_device(idx)
        nq = idx.shape[-1]
        flat = idx.reshape(-1, nq).contiguous().long()
        N = flat.shape[0]
        out = torch.empty((N, self.dim), dtype=torch.float32, device=idx.device)
        if N:
            cb = self.codebooks.detach()
            with torch.cuda.device(idx.device):
                _lib.check(lib.rvq_dequantize(_ptr(cb), _ptr(flat), N, N, 0, self.dim, 1, self.dim, first_stage, nq,
                                              self.K, None, 0, _ptr(out), _stream()), "rvq_dequantize")
        return out.reshape(*idx.shape[:-1], self.dim)

    # ------------------------------------------------------------------ wire format of the codes
    @property
    def code_bits(self) -> int:
        """Bits per code on the wire: ceil(log2 K) (bits per frame = n * code_bits, utils.py:137-147)."""
        return max(1, int(self.K - 1).bit_length())

    def pack_indices(self, idx: torch.Tensor) -> torch.Tensor:
        """idx (..., n) int64 -> uint8 (..., ceil(n * code_bits / 8)): codes of a frame LSB-first, byte-aligned frames."""
        lib = _lib.load()
        self._check_device(idx)
        nq = idx.shape[-1]
        flat = idx.reshape(-1, nq).contiguous().long()
        N = flat.shape[0]
        bpf = lib.rvq_packed_bytes_per_frame(nq, self.code_bits)
        out = torch.empty((N, bpf), dtype=torch.uint8, device=idx.device)
        with torch.cuda.device(idx.device):
            _lib.check(lib.rvq_pack_indices(_ptr(flat), N, nq, self.code_bits, _ptr(out), _stream()), "rvq_pack_indices")
        return out.reshape(*idx.shape[:-1], bpf)

    def unpack_indices(self, packed: torch.Tensor, n: Optional[int] = None) -> torch.Tensor:
        """Inverse of pack_indices: uint8 (..., bytes_per_frame) -> int64 (..., n)."""
        lib = _lib.load()
        self._check_device(packed)
        nq = self.num_quantizers if n is None else int(n)
        bpf = lib.rvq_packed_bytes_per_frame(nq, self.code_bits)
        if packed.dtype != torch.uint8 or packed.shape[-1] != bpf:
            raise ValueError(f"packed codes must be uint8 with {bpf} bytes per frame for n={nq}, K={self.K}")
        flat = packed.reshape(-1, bpf).contiguous()
        N = flat.shape[0]
        out = torch.empty((N, nq), dtype=torch.int64, device=packed.device)
        with torch.cuda.device(packed.device):
            _lib.check(lib.rvq_unpack_indices(_ptr(flat), N, nq, self.code_bits, _ptr(out), _stream()), "rvq_unpack_indices")
        return out.reshape(*packed.shape[:-1], nq)

    def decode_packed(self, packed: torch.Tensor, n: Optional[int] = None) -> torch.Tensor:
        """Wire bytes -> quantized latents (..., d): unpack + sum of code vectors (CausalVQAE.sample, vae.py:329-334)."""
        return self.dequantize(self.unpack_indices(packed, n))

    # ------------------------------------------------------------------ epoch-level API
    def get_stale_clusters(self):
        """Per stage, the number of codes whose EMA count (hits per call) is below ``vq_cutoff_freq`` - the codes the
        next update re-seeds (``training.py:435,461``; threshold semantics ASSUMED: Jukebox / lucidrains lineage)."""
        cnt = self.ema_count.cpu()
        return [int((cnt[q, :self.codebook_sizes[q]] < self.vq_cutoff_freq).sum())
                for q in range(self.num_quantizers)]

    def update_cutoff(self, new_cutoff=None, ratio=None):
        if new_cutoff is not None:
            self.vq_cutoff_freq = float(new_cutoff)
        if ratio is not None:
            self.vq_cutoff_freq *= float(ratio)

    def extra_repr(self):
        return (f"num_quantizers={self.num_quantizers}, dim={self.dim}, K={self.codebook_sizes}, "
                f"class={self.quantizer_class!r}, algo={self.algo!r}")


class HostEncoder:
    """End-to-end encode of HOST-resident frames: pinned host -> device -> RVQ -> codes back to pinned host.

    Frames are cut into chunks that ride ``n_buffers`` CUDA streams so that the host->device copy of chunk i+1 and
    the device->host copy of chunk i-1 overlap the kernel of chunk i.  This is the call a serving user makes when
    latents arrive from another process; ``bench.py`` times it as the ``e2e`` figure.

    The chunk size trades the pipeline's fill / drain (one chunk's copy in, one chunk's kernel and copy out per call)
    against per-copy overheads: 65536 frames measured best on C2 (``scripts/e2e_tune.py``, profiles/r2z_e2e_tune.log:
    103 M frames/s = 53 GB/s of host->device traffic for 1 M frames per call, against 55.6 GB/s for a bare 1 GiB copy).

    ``packed=True`` returns the codes in their wire format (``rvq_pack_indices``: ``ceil(n * code_bits / 8)`` bytes per
    frame instead of ``8 n``), which is what crosses PCIe back to the host.
    """

    def __init__(self, quantizer: ResidualQuantizer, chunk_frames: int = 1 << 16, n_buffers: int = 3,
                 packed: bool = False):
        self.q = quantizer
        self.chunk = int(chunk_frames)
        self.nbuf = int(n_buffers)
        self.packed = bool(packed)
        self._streams = None
        self._bufs = None

    def _setup(self, dev, nq):
        if self._streams is None:
            self._streams = [torch.cuda.Stream(device=dev) for _ in range(self.nbuf)]
            self._done = [torch.cuda.Event() for _ in range(self.nbuf)]
            d = self.q.dim
            self._bufs = [torch.empty(self.chunk, d, dtype=torch.float32, device=dev) for _ in range(self.nbuf)]
            self._wss = [torch.empty_like(self.q._workspace(dev)) for _ in range(self.nbuf)]

    def bytes_per_frame(self, n: Optional[int] = None) -> int:
        nq = self.q.num_quantizers if n is None else int(n)
        return _lib.load().rvq_packed_bytes_per_frame(nq, self.q.code_bits) if self.packed else 8 * nq

    @torch.no_grad()
    def encode(self, x_host: torch.Tensor, out_host: Optional[torch.Tensor] = None, n: Optional[int] = None):
        """x_host: pinned (N, d) fp32 on the CPU; returns pinned codes on the host, complete when the call returns:
        int64 (N, nq), or uint8 (N, bytes_per_frame) with ``packed=True``."""
        q = self.q
        nq = q.num_quantizers if n is None else int(n)
        dev = q.codebooks.device
        N = x_host.shape[0]
        if out_host is None:
            out_host = (torch.empty((N, self.bytes_per_frame(nq)), dtype=torch.uint8) if self.packed
                        else torch.empty((N, nq), dtype=torch.int64)).pin_memory()
        self._setup(dev, nq)
        q._prepared()                      # operands built on the current stream before the side streams start
        cur = torch.cuda.current_stream(dev)
        ready = torch.cuda.Event()
        ready.record(cur)
        used = set()
        for i, s0 in enumerate(range(0, N, self.chunk)):
            b = i % self.nbuf
            st, buf = self._streams[b], self._bufs[b]
            m = min(self.chunk, N - s0)
            st.wait_event(ready)
            with torch.cuda.stream(st):
                buf[:m].copy_(x_host[s0:s0 + m], non_blocking=True)
                _, idx, _, _ = q._encode(buf[:m], nq, False, ws=self._wss[b])
                out_host[s0:s0 + m].copy_(q.pack_indices(idx) if self.packed else idx, non_blocking=True)
                self._done[b].record(st)
            used.add(b)
        for b in used:
            cur.wait_event(self._done[b])     # stream order for device-side consumers ...
            self._done[b].synchronize()       # ... and the HOST may read the codes as soon as encode() returns
        return out_host
