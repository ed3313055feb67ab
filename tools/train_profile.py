"""One eager training step of the headline DLRM inside a cudaProfilerStart/Stop range (for `ncu --profile-from-start off`).

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv \
        python tools/train_profile.py [--batch 65536] [--opt adagrad]
Without ncu it prints the per-launch CUDA-event times of one step (kernel names from the call sequence).
"""
import argparse
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import models_b200 as mm  # noqa: E402
from models_b200 import datasets  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--opt", default="adagrad")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    mm.set_seed(1)
    schema = datasets.criteo_schema()
    model = mm.DLRMModel(schema, embedding_dim=64, bottom_block=mm.MLPBlock([128, 64]), top_block=mm.MLPBlock([128, 64, 32]),
                         embedding_options=mm.EmbeddingOptions(embeddings_initializers={"hash_seed": 4321}))
    model.build(dev)
    model.compile(optimizer=args.opt)
    b = datasets.generate_batch(schema, args.batch, seed=1, index_law="uniform", index_dtype=np.int32)
    feats, targets = datasets.split_targets(schema, b)
    x = {k: torch.from_numpy(v).to(dev) for k, v in feats.items()}
    y = torch.from_numpy(next(iter(targets.values()))).to(dev)
    tr = model.trainer(args.batch)
    for _ in range(2):
        tr.step(x, y)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    tr.step(x, y)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        tr.step(x, y)
    e1.record()
    torch.cuda.synchronize()
    print(f"eager step: {e0.elapsed_time(e1) / 5:.3f} ms, loss {tr.loss.item():.5f}")


if __name__ == "__main__":
    main()
