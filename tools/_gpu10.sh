mkdir -p gpurun_out
for cfg in "MM_IMMA_NBUF=2 MM_IMMA_WARPS=16" "MM_IMMA_NBUF=2 MM_IMMA_WARPS=8" "MM_IMMA_NBUF=4 MM_IMMA_WARPS=7" "MM_IMMA_NBUF=3 MM_IMMA_WARPS=10" "MM_IMMA_NBUF=4 MM_IMMA_WARPS=4"; do
  echo "== $cfg" >> gpurun_out/r2j_fused_peer.jsonl
  env $cfg timeout 300 python tools/probe/fused_peer.py >> gpurun_out/r2j_fused_peer.jsonl 2>> gpurun_out/r2j_fused_peer.err
done
echo done
