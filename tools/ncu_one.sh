k=$1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 2 -c 1 -f -o gpurun_out/r2_$k python tools/db_image_clk.py 64 > gpurun_out/ncu_$k.log 2>&1
ls -la gpurun_out/r2_$k.ncu-rep
