set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29540 tools/probe/peer_probe.py > gpurun_out/r2d_peer_probe.jsonl 2> gpurun_out/r2d_peer_probe.err; echo "probe rc=$?"
tail -5 gpurun_out/r2d_peer_probe.err
