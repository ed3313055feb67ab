mkdir -p gpurun_out
for cfg in "2 1" "32 1" "32 4" "32 16" "32 31"; do
  set -- $cfg
  PROBE_SHORT=1 PROBE_TABLE_GIB=$1 PROBE_SPAN_GIB=$2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29540 tools/probe/peer_probe.py >> gpurun_out/r2e_peer_probe_tlb.jsonl 2>> gpurun_out/r2e_peer_probe.err
done
echo done
