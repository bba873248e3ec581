cd $GRAFT_REPO_ROOT
export SCANN_B200_SCAN_TC=1
timeout 280 ncu --set full --clock-control none --import-source on -k regex:scan_tc_kernel -s 3 -c 1 -o gpurun_out/r02_scan_tc_v5 -f python bench.py --no-c5 --no-c3 --no-cpu-baseline --steps 1 > gpurun_out/r02_ncu_tc5.log 2>&1
echo rc=$?
