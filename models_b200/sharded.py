"""Row-sharded embedding tables across the GPUs of one box (BASELINE config 4).

Partitioning: row r of every table lives on rank `r % world` at local row `r // world` (balanced
under skew, keeps tiny tables from pinning to one GPU).  Dense / MLP / interaction weights are
replicated; the batch is sharded data-parallel.

Forward of the lookup on every rank (one process per GPU, `torch.distributed`):
  1. all-gather of the local index arrays -> GLOBAL indices (T, world*B_local) on every rank
     (4 B per feature per sample: tiny next to the 256-B rows);
  2. ONE kernel, `mm_shard_gather_push`: for the rows this rank owns it reads the row from its
     local shard and stores it straight into the destination rank's (B_local, F, D) stack through
     NVLink peer-mapped memory (torch symmetric memory) — gather and all-to-all fused, no pack /
     unpack buffers and no size exchange;
  3. a stream-ordered cross-rank barrier; everything downstream (interaction, MLPs) is replica-local.

The reference's counterpart is SOK's distributed variable + `sok.lookup_sparse`
(merlin/models/tf/distributed/embedding.py:75-84,144-148).  Host logic here (ownership maths,
index all-gather, stack layout) is backend-agnostic and covered by world-size-2 gloo tests on CPU;
the device step needs CUDA + peer access.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist

from . import _cabi, ops
from .core import create_variable
from .inputs import EmbeddingsBlock, _as_index


# ---- ownership maths (pure, also used by the tests) -------------------------------------------------
def owner_of(idx: torch.Tensor, world: int) -> torch.Tensor:
    """Rank that owns global row `idx` (Python modulo semantics: negative ids map into 0..world-1)."""
    return torch.remainder(idx, world)


def local_row(idx: torch.Tensor, world: int) -> torch.Tensor:
    return torch.div(idx, world, rounding_mode="floor")


def local_row_count(rows: int, rank: int, world: int) -> int:
    """Number of global rows r in [0, rows) with r % world == rank."""
    return (rows - rank + world - 1) // world if rows > rank else 0


def shard_of(full: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """The rows of a full table owned by `rank` (global rows rank, rank+world, ...)."""
    return full[rank::world].contiguous()


class ShardedEmbeddings:
    """Local shards of every table of an EmbeddingsBlock + the exchange that rebuilds, on every
    rank, the (B_local, F, D) feature stack of its own samples."""

    def __init__(self, embeddings: EmbeddingsBlock, group=None, device=None):
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.embeddings = embeddings
        self.device = device
        self.feature_names: List[str] = embeddings.feature_names
        dims = set(embeddings.output_dims().values())
        if len(dims) != 1:
            raise ValueError("sharded lookup needs one embedding dimension for all tables")
        self.D = dims.pop()
        self.shards: Dict[str, torch.Tensor] = {}
        self.global_rows: Dict[str, int] = {n: t.input_dim for n, t in embeddings.tables.items()}
        self._symm = None
        self._symm_key = None

    # ---- shard construction -------------------------------------------------------------------------
    def build(self, device) -> "ShardedEmbeddings":
        """Create the local shard of every table directly (never materialising a full table)."""
        self.device = device
        for name, table in self.embeddings.tables.items():
            if name in self.shards:
                continue
            lrows = local_row_count(table.input_dim, self.rank, self.world)
            init = table.embeddings_initializer
            w = torch.empty((max(lrows, 1), table.dim), dtype=torch.float32, device=device)
            if isinstance(init, dict) and "hash_seed" in init:
                _cabi.check(_cabi.load().mm_init_uniform_hash_rows(
                    w.data_ptr(), lrows, table.dim, init["hash_seed"] & (2**64 - 1), init.get("lo", -0.05),
                    init.get("hi", 0.05), self.rank, self.world, torch.cuda.current_stream().cuda_stream),
                    "mm_init_uniform_hash_rows")
            elif isinstance(init, (torch.Tensor,)) or hasattr(init, "shape"):
                full = torch.as_tensor(init, dtype=torch.float32)
                w = shard_of(full, self.rank, self.world).to(device)
            else:
                # seeded generators are rank-independent: build the full table row block by row block
                full = create_variable((table.input_dim, table.dim), init, device, f"{table.table_name}/embeddings")
                w = shard_of(full, self.rank, self.world)
                del full
            self.shards[name] = w
        return self

    def load_full_tables(self, full: Dict[str, torch.Tensor], device) -> "ShardedEmbeddings":
        """Shard explicitly given full tables (tests, checkpoints)."""
        self.device = device
        for name in self.embeddings.tables:
            self.shards[name] = shard_of(torch.as_tensor(full[name], dtype=torch.float32), self.rank, self.world).to(device)
        return self

    # ---- step 1: replicate the indices ----------------------------------------------------------------
    def gather_indices(self, local_inputs: Dict[str, torch.Tensor]) -> torch.Tensor:
        """(T, B_local) local -> (T, world*B_local) global, rank-major sample order."""
        idx = torch.stack([_as_index(local_inputs[f]).reshape(-1) for f in self.feature_names], dim=0).contiguous()
        if len({_as_index(local_inputs[f]).dtype for f in self.feature_names}) > 1:
            idx = idx.to(torch.int64)
        T, Bl = idx.shape
        out = torch.empty((self.world, T, Bl), dtype=idx.dtype, device=idx.device)
        dist.all_gather(list(out.unbind(0)), idx, group=self.group)  # works on NCCL and gloo alike
        return out.permute(1, 0, 2).reshape(T, self.world * Bl).contiguous()

    # ---- step 2 + 3: owner-computes push over peer memory, then barrier ------------------------------
    def _stack_buffer(self, B_local: int, width: int):
        """Symmetric (peer-mapped) fp32 buffer (B_local, width) + every rank's pointer to it."""
        key = (B_local, width)
        if self._symm_key != key:
            import torch.distributed._symmetric_memory as symm_mem

            buf = symm_mem.empty((B_local, width), dtype=torch.float32, device=self.device)
            hdl = symm_mem.rendezvous(buf, group=self.group)
            self._symm = (buf, hdl, [int(p) for p in hdl.buffer_ptrs])
            self._symm_key = key
        return self._symm

    def lookup_stack(self, local_inputs: Dict[str, torch.Tensor], slots: Dict[str, int], n_slots: int,
                     oob: Optional[torch.Tensor] = None) -> torch.Tensor:
        """(B_local, n_slots*D) stack with feature f at slot slots[f] (other slots untouched)."""
        if not self.shards:
            raise RuntimeError("ShardedEmbeddings.build(device) must be called first")
        g_idx = self.gather_indices(local_inputs)
        T, Bg = g_idx.shape
        Bl = Bg // self.world
        D = self.D
        buf, hdl, ptrs = self._stack_buffer(Bl, n_slots * D)
        arr = (_cabi.GatherTable * T)()
        for t, f in enumerate(self.feature_names):
            table = self.embeddings.feature_to_table[f]
            shard = self.shards[table.table_name]
            arr[t].weights = shard.data_ptr()
            arr[t].indices = g_idx[t].data_ptr()
            arr[t].rows = self.global_rows[table.table_name]
            arr[t].dim = D
            arr[t].out_col = slots[f] * D
        dst = (C.c_void_p * self.world)(*ptrs)
        hdl.barrier(channel=0)  # every rank is done reading the previous batch's stack
        _cabi.check(
            _cabi.load().mm_shard_gather_push(arr, T, ops._idx_dtype(g_idx, "indices"), Bg, Bl, D, self.rank, self.world,
                                              dst, n_slots * D, None if oob is None else oob.data_ptr(),
                                              torch.cuda.current_stream().cuda_stream),
            "mm_shard_gather_push")
        hdl.barrier(channel=1)  # all pushes have landed
        return buf


def shard_model(model, group=None):
    """Row-shard the embedding tables of a DLRM model over `group` (call before the first forward;
    every rank then holds 1/world of each table).  Returns the model."""
    from .blocks import DLRM

    body = getattr(model, "body", model)
    if not isinstance(body, DLRM):
        raise NotImplementedError("shard_model supports DLRM bodies")
    if any(t.table is not None for t in body.embeddings.tables.values()):
        raise RuntimeError("shard_model must be called before the tables are built")
    body.sharded = ShardedEmbeddings(body.embeddings, group)
    return model
