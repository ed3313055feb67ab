#!/usr/bin/env python
"""Per-kernel microbenchmarks (CUDA events, L2-cold inputs: tables >> L2, rotating buffers).

    python tools/microbench.py [--only gather,interact,fused,dense,scores] [--iters 20]

Prints one JSON line per kernel with achieved algorithmic GB/s (or TFLOP/s) and the fraction of
the measured peak (MEASURED_PEAKS.json).  Used to produce profiles/*.md; never under a profiler.
"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import models_b200 as mm  # noqa: E402
from models_b200 import datasets, ops  # noqa: E402


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d["hbm_gbs"], d["bf16_tflops"]
    return 6650.0, 1590.0


def timeit(fn, iters, warmup=3):
    for _ in range(warmup):
        fn(0)
    torch.cuda.synchronize()
    evs = []
    for i in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn(i)
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    t = sorted(x.elapsed_time(y) for x, y in evs)
    return float(np.mean(t)), float(t[0])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="gather,interact,fused,dense,scores")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--law", default="uniform")
    args = ap.parse_args()
    only = set(args.only.split(","))
    dev = torch.device("cuda", 0)
    hbm, tf = peaks()
    B, T, D = args.batch, 26, 64
    F = T + 1
    res = []

    def report(name, ms_mean, ms_min, bytes_=None, flops=None, **extra):
        r = {"kernel": name, "ms_mean": ms_mean, "ms_min": ms_min, **extra}
        if bytes_ is not None:
            r["algorithmic_bytes"] = bytes_
            r["GBps"] = bytes_ / (ms_mean * 1e-3) / 1e9
            r["frac_hbm"] = r["GBps"] / hbm
        if flops is not None:
            r["TFLOPs"] = flops / (ms_mean * 1e-3) / 1e12
            r["frac_bf16_peak"] = r["TFLOPs"] / tf
        print(json.dumps(r), flush=True)
        res.append(r)

    if only & {"gather", "interact", "fused"}:
        schema = datasets.criteo_schema()
        cat = schema.select_by_tag(mm.Tags.CATEGORICAL)
        emb = mm.Embeddings(cat, dim=D, embeddings_initializer={"hash_seed": 4321})
        emb.build(dev)
        names = emb.feature_names
        slots = {n: i for i, n in enumerate(sorted(names + ["bottom_block"]))}
        nb = 4
        idx = []
        for i in range(nb):
            b = datasets.generate_batch(cat, B, seed=100 + i, index_law=args.law)
            idx.append([torch.from_numpy(b[n]).to(dev) for n in names])
        tables = [emb.feature_to_table[n].table for n in names]
        stack = [torch.empty((B, F * D), dtype=torch.float32, device=dev) for _ in range(2)]
        bottom = torch.randn((B, D), device=dev)
        out = [torch.empty((B, D + F * (F - 1) // 2), dtype=torch.float32, device=dev) for _ in range(2)]
        if "gather" in only:
            m, mn = timeit(lambda i: ops.gather_multi(tables, idx[i % nb], [slots[n] * D for n in names], stack[i % 2]), args.iters)
            report("mm_gather_multi (26 tables, D=64, int32)", m, mn, bytes_=B * (2 * T * D * 4 + T * 4), law=args.law)
        if "interact" in only:
            x = torch.randn((B, F, D), device=dev)
            m, mn = timeit(lambda i: ops.dot_interaction(x, out[i % 2], prefix=bottom), args.iters)
            report("mm_dot_interaction (F=27, D=64, +prefix)", m, mn, bytes_=B * (F * D * 4 + (D + 351) * 4),
                   flops=B * 351 * 64 * 2)
        if "fused" in only:
            m, mn = timeit(lambda i: ops.dlrm_gather_interact(tables, idx[i % nb], [slots[n] for n in names], D, bottom,
                                                              slots["bottom_block"], out[i % 2]), args.iters)
            report("mm_dlrm_gather_interact", m, mn, bytes_=B * (T * D * 4 + T * 4 + D * 4 + (D + 351) * 4), law=args.law)

    if "dense" in only:
        for (K, N) in [(13, 128), (128, 64), (415, 128), (128, 64), (64, 32), (32, 1), (1037, 1037), (1024, 1024)]:
            x = torch.randn((B, K), device=dev)
            W = torch.randn((K, N), device=dev) * 0.05
            b = torch.zeros(N, device=dev)
            o = torch.empty((B, N), device=dev)
            m, mn = timeit(lambda i: ops.dense_fp32(x, W, b, "relu", o), max(5, args.iters // 2))
            report(f"mm_dense_fp32 {K}->{N}", m, mn, bytes_=B * (K + N) * 4, flops=2.0 * B * K * N)

    if "scores" in only:
        Bq, Dq = 16384, 64
        q = torch.randn((Bq, Dq), device=dev)
        it = torch.randn((Bq, Dq), device=dev)
        ids = torch.randint(0, 10_000_000, (Bq,), device=dev, dtype=torch.int64)
        o = torch.empty((Bq, Bq + 4), device=dev)[:, 3:]  # negatives 16-B aligned (as retrieval._score allocates)
        for tc in (False, True):
            m, mn = timeit(lambda i: ops.inbatch_scores(q, it, it, o, pos_ids=ids, neg_ids=ids, tensor_cores=tc), max(5, args.iters // 2))
            report(f"mm_inbatch_scores 16384x16384 D=64 ({'tcgen05 split-bf16 incl. operand split' if tc else 'fp32 SIMT'})", m, mn,
                   bytes_=Bq * (Bq + 1) * 4, flops=2.0 * Bq * Bq * Dq)


if __name__ == "__main__":
    main()
