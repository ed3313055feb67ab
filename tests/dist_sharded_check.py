"""Multi-GPU check of the row-sharded DLRM forward (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/dist_sharded_check.py

Every rank builds the SAME full tables (seeded), rank r keeps rows r, r+world, ... ; the sharded
forward of the rank's local batch must equal the unsharded forward of the same batch, and the
rebuilt (B_local, F, D) stack must be bit-identical to a local gather."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import models_b200 as mm  # noqa: E402
from models_b200 import datasets  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    schema = datasets.criteo_schema({k: min(v, 3000) for k, v in datasets.CRITEO_MAX.items()})

    def make():
        mm.set_seed(99)
        return mm.DLRMModel(schema, embedding_dim=64, bottom_block=mm.MLPBlock([32, 64]), top_block=mm.MLPBlock([64, 16]))

    full = make()
    full.build(dev)
    tables = {n: t.embeddings.clone() for n, t in full.body.embeddings.tables.items()}
    ok = True
    # two placements: every table row-sharded; and big tables sharded + small tables replicated
    for replicate_below in (0, 2000):
        sharded = mm.shard_model(make(), replicate_below_rows=replicate_below)
        se = sharded.body.sharded
        se.load_full_tables(tables, dev)
        sharded.build(dev)
        for name, t in tables.items():
            want_rows = t[rank::world] if se.is_sharded(name) else t
            # a rank that owns no row of a tiny table (rows < world) still holds a 1-row placeholder shard
            assert torch.equal(se.shards[name][: want_rows.shape[0]], want_rows), name
        batches = []
        for step in range(3):
            B = 512
            b, _ = datasets.split_targets(schema, datasets.generate_batch(schema, B, seed=100 * rank + step, index_law="uniform",
                                                                          index_dtype=np.int32))
            if step == 2:
                b["C2"] = b["C2"].copy()
                b["C2"][:] = b["C2"][0]  # skew: every sample hits the same row (one owner)
            batches.append(b)
            d = {k: torch.from_numpy(v).to(dev) for k, v in b.items()}
            want = full(d)
            got = sharded(d)  # product path: lookup over NVLink inside the interaction kernel
            ok &= bool(torch.equal(want, got))
            # NCCL baseline (ids all-gather + local gather + one variable-size all-to-all + scatter): same rows
            from models_b200.sharded import lookup_stack_nccl

            slots_n = sharded.body.slots()
            st = lookup_stack_nccl(se, d, slots_n, len(slots_n))
            ref_n = torch.zeros_like(st)
            full.body.embeddings.lookup_all_into(d, ref_n, {f: slots_n[f] * 64 for f in full.body.embeddings.feature_names})
            cols_n = [c for f in full.body.embeddings.feature_names for c in range(slots_n[f] * 64, slots_n[f] * 64 + 64)]
            ok &= bool(torch.equal(st[:, cols_n], ref_n[:, cols_n]))
            if replicate_below == 0:
                # staged protocol (index all-gather + push + barrier): stack bit-exactness
                slots = sharded.body.slots()
                stack = se.lookup_stack(d, slots, len(slots)).clone()
                ref = torch.zeros_like(stack)
                full.body.embeddings.lookup_all_into(d, ref, {f: slots[f] * 64 for f in full.body.embeddings.feature_names})
                cols = [c for f in full.body.embeddings.feature_names for c in range(slots[f] * 64, slots[f] * 64 + 64)]
                ok &= bool(torch.equal(stack[:, cols], ref[:, cols]))
        # the sharded forward captures into a CUDA graph like the replicated one (no collective inside);
        # host batches with packed ids go through the same graph
        hb = [mm.HostBatch.like(b, sharded.input_columns(), id_bytes=sharded.id_bytes()) for b in batches]
        cf = sharded.compile(hb[0])
        for i, b in enumerate(batches):
            d = {k: torch.from_numpy(v).to(dev) for k, v in b.items()}
            ok &= bool(torch.equal(cf(hb[i]).to(dev), full(d)))
        dist.barrier()
        del cf
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("SHARDED_OK" if int(flag.item()) == 1 else "SHARDED_MISMATCH", "world", world)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
