"""Host logic of the row-sharded lookup on CPU with world_size = 2 over gloo: ownership maths,
shard construction, the index all-gather, and — emulating the device push with the NumPy oracle
and a gloo exchange — that owner-computes + exchange rebuilds each rank's (B_local, F, D) stack."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import models_b200 as mm
from models_b200 import datasets, sharded


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_ownership_maths():
    idx = torch.tensor([0, 1, 2, 3, 7, 8, 9, -1, -2])
    assert sharded.owner_of(idx, 4).tolist() == [0, 1, 2, 3, 3, 0, 1, 3, 2]
    assert sharded.local_row(idx, 4).tolist()[:7] == [0, 0, 0, 0, 1, 2, 2]
    for rows in (1, 3, 4, 10, 4097):
        for world in (1, 2, 3, 8):
            counts = [sharded.local_row_count(rows, r, world) for r in range(world)]
            assert sum(counts) == rows
            full = torch.arange(rows * 2, dtype=torch.float32).reshape(rows, 2)
            for r in range(world):
                sh = sharded.shard_of(full, r, world)
                assert sh.shape[0] == counts[r]
                if counts[r]:
                    assert torch.equal(sh[:, 0] / 2, torch.arange(r, rows, world, dtype=torch.float32))


def _worker(rank, world, port, result):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle

        schema = datasets.criteo_schema({k: min(v, 97) for k, v in datasets.CRITEO_MAX.items()})
        cat = schema.select_by_tag(mm.Tags.CATEGORICAL)
        emb = mm.Embeddings(cat, dim=8)
        se = sharded.ShardedEmbeddings(emb, replicate_below_rows=0)
        rng = np.random.default_rng(5)
        full = {n: rng.standard_normal((t.input_dim, 8)).astype(np.float32) for n, t in emb.tables.items()}
        se.load_full_tables({n: torch.from_numpy(v) for n, v in full.items()}, torch.device("cpu"))
        Bl = 16
        batch = datasets.generate_batch(cat, Bl, seed=40 + rank, index_law="uniform")
        local = {k: torch.from_numpy(v) for k, v in batch.items()}
        g_idx = se.gather_indices(local)  # (T, world*Bl) over gloo
        T = len(se.feature_names)
        assert tuple(g_idx.shape) == (T, world * Bl)
        for t, f in enumerate(se.feature_names):
            assert np.array_equal(g_idx[t, rank * Bl:(rank + 1) * Bl].numpy(), batch[f])
        # emulate mm_shard_gather_push: this rank contributes the rows it owns for EVERY global sample;
        # exchanged with a (dense) gloo all-to-all and summed, exactly one contribution per row is non-zero
        contrib = np.zeros((world, Bl, T, 8), dtype=np.float32)
        for t, f in enumerate(se.feature_names):
            name = emb.feature_to_table[f].table_name
            shard = se.shards[name].numpy()
            idx = g_idx[t].numpy().astype(np.int64)
            mine = sharded.owner_of(torch.from_numpy(idx), world).numpy() == rank
            lrow = sharded.local_row(torch.from_numpy(idx), world).numpy()
            rows = np.zeros((world * Bl, 8), dtype=np.float32)
            rows[mine] = oracle.embedding_lookup(shard, lrow[mine])
            contrib[:, :, t, :] = rows.reshape(world, Bl, 8)
        # gloo has no all-to-all: every rank gathers all contributions and keeps the slice addressed to it
        mine_t = torch.from_numpy(contrib)
        gathered = [torch.empty_like(mine_t) for _ in range(world)]
        dist.all_gather(gathered, mine_t)
        stack = sum(gsrc[rank].numpy() for gsrc in gathered)
        want = np.stack([oracle.embedding_lookup(full[emb.feature_to_table[f].table_name], batch[f]) for f in se.feature_names], axis=1)
        result[rank] = bool(np.array_equal(stack, want))
    finally:
        dist.destroy_process_group()


def test_sharded_protocol_world2_gloo():
    world = 2
    mgr = mp.Manager()
    result = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), result), nprocs=world, join=True)
    assert dict(result) == {0: True, 1: True}


def _loader_worker(rank, world, port, result):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 1003
        data = {"row": np.arange(n, dtype=np.int64), "x": np.arange(n, dtype=np.float32) * 0.5, "y": (np.arange(n) % 2).astype(np.int64)}
        loader = mm.Loader(data, batch_size=100, shuffle=True, seed_fn=lambda: 99, label_names=["y"], device="cpu",
                           global_size=dist.get_world_size(), global_rank=dist.get_rank())
        rows = torch.cat([b[0]["row"] for b in loader]).to(torch.int64)
        for inputs, targets in loader:  # features and targets stay aligned with their rows after the sharded shuffle
            assert torch.equal(inputs["x"], inputs["row"].to(torch.float32) * 0.5)
            assert torch.equal(targets, inputs["row"] % 2)
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([rows.numel()]))
        pad = torch.full((int(max(s.item() for s in sizes)),), -1, dtype=torch.int64)
        pad[: rows.numel()] = rows
        gathered = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(gathered, pad)
        allrows = torch.cat(gathered)
        allrows = allrows[allrows >= 0]
        ok = allrows.numel() == n and torch.equal(torch.sort(allrows).values, torch.arange(n))
        ok = ok and abs(rows.numel() - n / world) <= 1 and len(loader) == (rows.numel() + 99) // 100
        result[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_loader_shards_partition_the_dataset_across_ranks():
    """One Loader per process (as one rank per GPU): the ranks' shuffled shards are disjoint, equally sized (±1) and
    together cover every row exactly once — checked with a gloo all_gather over world_size 2."""
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    result = mgr.dict()
    mp.spawn(_loader_worker, args=(world, port, result), nprocs=world, join=True)
    assert dict(result) == {0: True, 1: True}


def _gather_worker(rank, world, port, result):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from models_b200.train import gather_slices

        T, B, D = 3, 5, 4
        ids = (torch.arange(T * B, dtype=torch.int32).reshape(T, B) + 100 * rank)
        sl = torch.arange(T * B * D, dtype=torch.float32).reshape(T, B, D) + 1000.0 * rank
        all_ids, all_sl = gather_slices(ids, sl, dist.group.WORLD)
        ok = tuple(all_ids.shape) == (T, world * B) and tuple(all_sl.shape) == (T, world * B, D)
        for r in range(world):  # table-major, then rank, then the rank's samples in order: one IndexedSlices per table
            ok &= bool(torch.equal(all_ids[:, r * B:(r + 1) * B], torch.arange(T * B, dtype=torch.int32).reshape(T, B) + 100 * r))
            ok &= bool(torch.equal(all_sl[:, r * B:(r + 1) * B], torch.arange(T * B * D, dtype=torch.float32).reshape(T, B, D) + 1000.0 * r))
        # the dense arena is averaged with one all-reduce
        g = torch.full((7,), float(rank + 1))
        dist.all_reduce(g)
        ok &= bool(torch.equal(g / world, torch.full((7,), (world + 1) / 2)))
        result[rank] = ok
    finally:
        dist.destroy_process_group()


def test_training_gradient_exchange_world2_gloo():
    """Host side of the data-parallel training step (models_b200/train.py): layout of the all-gathered IndexedSlices."""
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    result = mgr.dict()
    mp.spawn(_gather_worker, args=(world, port, result), nprocs=world, join=True)
    assert all(result[r] for r in range(world)), dict(result)
