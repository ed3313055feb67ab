export MM_INTERACT_TC=1
for d in 0 1 2 4 3 7; do echo "DBG=$d"; MM_ITC_DBG=$d timeout 200 python tools/microbench.py --only fused 2>&1 | tail -1 | cut -c1-110; done
