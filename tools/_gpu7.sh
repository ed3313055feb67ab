mkdir -p gpurun_out
PROBE_MIXED=1 PROBE_TABLE_GIB=16 PROBE_SPAN_GIB=15 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29540 tools/probe/peer_probe.py > gpurun_out/r2g_peer_probe_mixed.jsonl 2> gpurun_out/r2g_peer_probe.err
echo rc=$?
