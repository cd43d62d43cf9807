for wl in cfg3 davis; do
ncu --set full --import-source on --clock-control none --kernel-name regex:k_edge_pass_v2 --launch-skip 6 --launch-count 1 -f -o gpurun_out/k1v2_$wl python tools/stage_times.py $wl > gpurun_out/ncu_k1_$wl.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
