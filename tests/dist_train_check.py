"""Data-parallel training step on >= 2 GPUs (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
        tests/dist_train_check.py

Every rank holds the same model and trains on ITS shard of a global batch with a process group (dense gradient arena
all-reduced, embedding IndexedSlices all-gathered and divided by the world size — Horovod's DistributedOptimizer,
merlin/models/tf/models/base.py:476-508).  Because the loss is a mean over equally sized shards, the result must equal a
single-GPU step on the concatenated global batch; rank 0 trains that twin and compares every variable after 3 steps.
"""
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


def variables(model):
    out = [t.embeddings for _, t in sorted(model.body.embeddings.tables.items())]
    for blk in (model.body.bottom_block, model.body.top_block):
        for l in blk.dense_layers:
            out += [l.kernel, l.bias]
    return out + [model.prediction.to_call.kernel, model.prediction.to_call.bias]


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    schema = datasets.criteo_schema({k: min(v, 500) for k, v in datasets.CRITEO_MAX.items()})

    def make():
        mm.set_seed(21)
        m = mm.DLRMModel(schema, embedding_dim=32, bottom_block=mm.MLPBlock([64, 32]), top_block=mm.MLPBlock([64, 16]))
        m.build(dev)
        return m

    B = 256
    ok = True
    for opt in ("sgd", "adagrad", "adam"):
        def optimizer():
            return {"sgd": mm.SGD(0.5), "adagrad": mm.Adagrad(0.05), "adam": mm.Adam(0.01, epsilon=1e-6)}[opt]

        dp = make()
        dp.compile(optimizer=optimizer())
        tr = dp.trainer(B, group=dist.group.WORLD)
        before = [v.clone() for v in variables(dp)]
        twin = None
        if rank == 0:
            twin = make()
            twin.compile(optimizer=optimizer())
            twin_tr = twin.trainer(B * world)
        for step in range(3):
            g = datasets.generate_batch(schema, B * world, seed=500 + step, index_law="uniform", index_dtype=np.int32)
            feats, targets = datasets.split_targets(schema, g)
            y = next(iter(targets.values()))
            mine = {k: torch.from_numpy(v[rank * B:(rank + 1) * B].copy()).to(dev) for k, v in feats.items()}
            loss = tr.step(mine, torch.from_numpy(y[rank * B:(rank + 1) * B].copy()).to(dev)).clone()
            dist.all_reduce(loss)
            if rank == 0:
                lt = twin_tr.step({k: torch.from_numpy(v).to(dev) for k, v in feats.items()}, torch.from_numpy(y).to(dev))
                ok &= bool(abs(loss.item() / world - lt.item()) < 1e-5 * max(1.0, abs(lt.item())))
        # every rank ends with the same variables
        for v in variables(dp):
            ref = v.clone()
            dist.broadcast(ref, 0)
            ok &= bool(torch.allclose(v, ref, rtol=0, atol=1e-6))
        if rank == 0:
            for i, (a, b, b0) in enumerate(zip(variables(dp), variables(twin), before)):
                ua, ub = (a - b0).double(), (b - b0).double()
                err = float((ua - ub).norm() / ub.norm().clamp_min(1e-30))
                if not err < 2e-3:
                    print(f"{opt}: variable {i}: update differs from the single-GPU twin: {err:.3e}")
                    ok = False
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("TRAIN_DP_OK world", world if int(flag.item()) else "FAILED")
    dist.destroy_process_group()
    return 0 if int(flag.item()) else 1


if __name__ == "__main__":
    sys.exit(main())
