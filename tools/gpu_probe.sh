# one GPU visit: stage times of the K3 / K1 A/B builds, then an ncu capture of the default build's K3 and K1 (one launch each)
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 || { echo "SMOKE FAILED OR HUNG"; exit 1; }
echo "== default"; timeout 200 python tools/stage_probe.py 128 stats 2>&1 | tail -2
for v in "$@"; do echo "== $v"; MCS_B200_LIB=$PWD/multicol_slam_b200/libmcs_b200_$v.so timeout 200 python tools/stage_probe.py 128 2>&1 | tail -1; done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:describe_kernel -c 1 -o gpurun_out/r2_k3 -f python tools/stage_probe.py 32 > gpurun_out/ncu_k3.log 2>&1; tail -2 gpurun_out/ncu_k3.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:pyr_fast_kernel -c 8 -o gpurun_out/r2_k1 -f python tools/stage_probe.py 32 > gpurun_out/ncu_k1.log 2>&1; tail -2 gpurun_out/ncu_k1.log
