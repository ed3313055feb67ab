export MM_INTERACT_TC=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:interact_tc -c 1 -o gpurun_out/itc_v4 python tools/run_kernel.py fused > gpurun_out/itc_ncu.log 2>&1
tail -2 gpurun_out/itc_ncu.log
