"""Bulk embedding extraction — the caller side of the hot path (host-side mirror of what
/root/reference/scripts/extract.py:63-131 does per batch: tokens to the device, `model(toks, repr_layers)`, results
back to the host, per-sequence slicing / mean pooling).

`BulkEmbedder.embed(tokens)` takes HOST tokens and returns HOST results; internally the batch is cut into
micro-batches so that the device->host copy of micro-batch i (pinned memory, copy stream) overlaps the compute of
micro-batch i+1.  This is the call bench.py's `e2e` number times.

`shard_range` / `ShardedEmbedder` split a batch of independent sequences over the ranks of one node (one process per
GPU, torch.distributed) and all-gather the per-sequence mean representations with a single collective
(SURVEY §8e); the per-token representations stay on the rank that computed them.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Sequence, Tuple

import torch

from . import _lib
from .model import ESM2, _ptr, _stream


def mean_pool(x: torch.Tensor, lengths: torch.Tensor) -> torch.Tensor:
    """x fp32 [B,T,E] (cuda), lengths int32 [B] (cuda) -> [B,E]: mean over residues 1..len (extract.py:116-119)."""
    B, T, E = x.shape
    out = torch.empty((B, E), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().esmb200_mean_pool(_ptr(x), _ptr(lengths), _ptr(out), B, T, E, _stream()))
    return out


def residue_lengths(tokens: torch.Tensor, alphabet) -> torch.Tensor:
    """number of residues per row = non-pad tokens minus <cls>/<eos> (extract.py:109 uses len(strs[i]))."""
    n = tokens.ne(alphabet.padding_idx).sum(-1)
    n = n - int(alphabet.prepend_bos) - int(alphabet.append_eos)
    return n.to(torch.int32)


class BulkEmbedder:
    """Host-to-host embedding of token batches with copy/compute overlap.

    include: any of "mean" ([B,E]), "per_tok" ([B,T,E]), "bos" ([B,E]) — the keys of scripts/extract.py:41-46.
    """

    def __init__(self, model: ESM2, repr_layer: Optional[int] = None, include: Sequence[str] = ("mean",),
                 micro_batch: int = 32):
        self.model = model
        self.repr_layer = model.num_layers if repr_layer is None else repr_layer
        self.include = tuple(include)
        self.micro_batch = micro_batch
        self.device = next(model.parameters()).device
        if self.device.type != "cuda":
            raise _lib.Esmb200Error("BulkEmbedder needs the model on a CUDA device (no CPU fallback)")
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._host: Dict[Tuple, torch.Tensor] = {}
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def _pinned(self, name: str, shape, dtype) -> torch.Tensor:
        key = (name, tuple(shape), dtype)
        buf = self._host.get(key)
        if buf is None:
            buf = torch.empty(shape, dtype=dtype, pin_memory=True)
            self._host = {k: v for k, v in self._host.items() if k[0] != name}
            self._host[key] = buf
        return buf

    @torch.no_grad()
    def embed(self, tokens: torch.Tensor) -> Dict[str, torch.Tensor]:
        """tokens: int64 [B,T] on the HOST (pinned for asynchronous H2D). Returns host tensors (pinned, reused
        between calls: clone what must outlive the next call)."""
        assert tokens.device.type == "cpu" and tokens.dtype == torch.int64 and tokens.ndim == 2
        B, T = tokens.shape
        E = self.model.embed_dim
        res: Dict[str, torch.Tensor] = {}
        if "mean" in self.include:
            res["mean"] = self._pinned("mean", (B, E), torch.float32)
        if "bos" in self.include:
            res["bos"] = self._pinned("bos", (B, E), torch.float32)
        if "per_tok" in self.include:
            res["per_tok"] = self._pinned("per_tok", (B, T, E), torch.float32)
        self.h2d_bytes = tokens.numel() * 8
        self.d2h_bytes = sum(v.numel() * 4 for v in res.values())
        compute = torch.cuda.current_stream(self.device)
        mb = self.micro_batch
        pending = []
        with torch.cuda.device(self.device):
            for s in range(0, B, mb):
                e = min(B, s + mb)
                tok_d = tokens[s:e].to(self.device, non_blocking=True)
                out = self.model(tok_d, repr_layers=[self.repr_layer])["representations"][self.repr_layer]
                dev = {}
                if "mean" in res:
                    dev["mean"] = mean_pool(out, residue_lengths(tok_d, self.model.alphabet))
                if "bos" in res:
                    dev["bos"] = out[:, 0].contiguous()
                if "per_tok" in res:
                    dev["per_tok"] = out
                done = torch.cuda.Event()
                done.record(compute)
                self.copy_stream.wait_event(done)
                with torch.cuda.stream(self.copy_stream):
                    for k, v in dev.items():
                        res[k][s:e].copy_(v, non_blocking=True)
                        v.record_stream(self.copy_stream)
                pending.append(dev)
            self.copy_stream.synchronize()
        return res


# ------------------------------------------------------------------------------------------------------------------
# multi-GPU: shard independent sequences over ranks, one all-gather of the per-sequence representations
# ------------------------------------------------------------------------------------------------------------------
def shard_range(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced split: the first n_items % world_size ranks get one extra item."""
    base, rem = divmod(n_items, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def all_gather_rows(local: torch.Tensor, n_items: int, group=None) -> torch.Tensor:
    """All-gather row blocks produced under `shard_range` into one [n_items, ...] tensor on every rank with a single
    collective (NCCL on GPUs, gloo in the CPU tests).  Uneven shards are padded to the largest shard."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    max_rows = (n_items + world - 1) // world
    s, e = shard_range(n_items, world, rank)
    assert local.shape[0] == e - s
    pad = torch.zeros((max_rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: e - s] = local
    gathered = torch.empty((world * max_rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, pad, group=group)
    if n_items % world == 0:
        return gathered
    parts = []
    for r in range(world):
        rs, re = shard_range(n_items, world, r)
        parts.append(gathered[r * max_rows: r * max_rows + (re - rs)])
    return torch.cat(parts, 0)


class ShardedEmbedder:
    """One process per GPU: rank r embeds sequences shard_range(B, W, r) of every batch and all ranks end up with the
    [B,E] mean representations (the reference has no multi-GPU path: scripts/extract.py:70-72 uses one device)."""

    def __init__(self, model: ESM2, repr_layer: Optional[int] = None, group=None):
        self.model = model
        self.repr_layer = model.num_layers if repr_layer is None else repr_layer
        self.group = group

    @torch.no_grad()
    def embed_mean(self, tokens: torch.Tensor, keep_per_tok: bool = False):
        """tokens int64 [B,T] (same on every rank, already on this rank's device). Returns ([B,E] on every rank,
        local per-token representations or None)."""
        import torch.distributed as dist
        B = tokens.shape[0]
        s, e = shard_range(B, dist.get_world_size(self.group), dist.get_rank(self.group))
        local_tok = tokens[s:e].contiguous()
        out = self.model(local_tok, repr_layers=[self.repr_layer])["representations"][self.repr_layer]
        local_mean = mean_pool(out, residue_lengths(local_tok, self.model.alphabet))
        return all_gather_rows(local_mean, B, self.group), (out if keep_per_tok else None)
