cd $GRAFT_REPO_ROOT
timeout 280 ncu --set full --clock-control none --import-source on -k regex:scan_tc_kernel -s 3 -c 1 -o gpurun_out/r02_scan_tc_v2 -f python bench.py --no-c5 --no-c3 --no-cpu-baseline --steps 1 > gpurun_out/r02_ncu_tc2.log 2>&1
echo rc=$?
