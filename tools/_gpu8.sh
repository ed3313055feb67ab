mkdir -p gpurun_out
timeout 300 python tools/probe/fused_peer.py > gpurun_out/r2h_fused_peer.jsonl 2> gpurun_out/r2h_fused_peer.err; echo rc=$?
timeout 400 ncu --set full --clock-control none --import-source on -k regex:interact_v2 -s 3 -c 1 -o gpurun_out/r2h_fused_peer -f python tools/probe/fused_peer.py --only-idle --iters 3 > gpurun_out/r2h_ncu.log 2>&1; echo ncu rc=$?
