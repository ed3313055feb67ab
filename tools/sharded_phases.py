#!/usr/bin/env python
"""Phase timing of the row-sharded DLRM step (torchrun, one rank per GPU): where do the milliseconds go?

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/sharded_phases.py
"""
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import models_b200 as mm  # noqa: E402
from models_b200 import datasets  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
B = int(os.environ.get("MM_BATCH", "65536"))
schema = datasets.criteo_tb_schema()
model = mm.DLRMModel(schema, embedding_dim=64, bottom_block=mm.MLPBlock([128, 64]), top_block=mm.MLPBlock([128, 64, 32]),
                     embedding_options=mm.EmbeddingOptions(embeddings_initializers={"hash_seed": 4321}))
mm.shard_model(model)
model.build(dev)
b, _ = datasets.split_targets(schema, datasets.generate_batch(schema, B, seed=4000 + rank, index_law="uniform"))
d = {k: torch.from_numpy(v).to(dev) for k, v in b.items()}
sh = model.body.sharded
events = {}


def timed(name, fn):
    def wrap(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn(*a, **k)
        e1.record()
        events.setdefault(name, []).append((e0, e1))
        return r
    return wrap


sh.gather_indices = timed("gather_indices (stack + NCCL all_gather + permute)", sh.gather_indices)
sh.lookup_stack = timed("lookup_stack (indices + barrier + push + barrier)", sh.lookup_stack)
model.body.bottom_forward = timed("bottom path", model.body.bottom_forward)
for _ in range(5):
    model(d)
torch.cuda.synchronize()
dist.barrier(device_ids=[local])
events.clear()
N = 20
t0 = time.perf_counter()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(N):
    model(d)
e1.record()
host = (time.perf_counter() - t0) / N * 1e3
torch.cuda.synchronize()
if rank == 0:
    print(f"world {world}  B/GPU {B}: step {e0.elapsed_time(e1) / N:.3f} ms on the device, {host:.3f} ms of host enqueue time per step")
    for k, v in events.items():
        ms = sorted(a.elapsed_time(b) for a, b in v)
        print(f"  {k}: median {ms[len(ms) // 2]:.3f} ms")
# the same step captured into a CUDA graph (NCCL all-gather, symmetric-memory barriers and the push kernel inside)
if os.environ.get("MM_SHARDED_GRAPH", "1") == "1":
    try:
        ref = model(d).clone()
        cf = model.compile(b)
        hb = mm.HostBatch.like(b, model.input_columns())
        for _ in range(3):
            cf.replay()
        torch.cuda.synchronize()
        dist.barrier(device_ids=[local])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(N):
            cf.replay()
        e1.record()
        host = (time.perf_counter() - t0) / N * 1e3
        torch.cuda.synchronize()
        same = bool(torch.equal(cf.output, ref))
        if rank == 0:
            print(f"graph replay: step {e0.elapsed_time(e1) / N:.3f} ms on the device, {host:.3f} ms host; {cf.launches_per_replay} launches; "
                  f"output == eager: {same}")
    except Exception as e:  # noqa: BLE001
        print(f"rank {rank}: graph capture of the sharded step failed: {type(e).__name__}: {e}")
dist.destroy_process_group()
