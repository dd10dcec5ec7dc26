# ncu --set full capture of the expansion kernels of one PSE / PAN bench step (all four tile classes of one call)
w=$1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ex_expand_kernel --launch-skip 8 -c 4 -f -o gpurun_out/r2_${w}_expand python bench.py --workload $w --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_${w}_expand.log 2>&1
ls -la gpurun_out/r2_${w}_expand.ncu-rep
