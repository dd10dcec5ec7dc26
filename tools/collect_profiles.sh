set -x
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_final_default.json 2> gpurun_out/r2_final_default.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_final_ref_db.json 2> gpurun_out/r2_final_ref_db.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_db_launches.csv python bench.py --steps 2 --warmup 1 --headline-only --no-cpu > gpurun_out/ncu_launches.log 2>&1
for k in db_scan_kernel db_image_kernel db_geometry_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 2 -c 1 -f -o gpurun_out/r2_$k python tools/db_image_clk.py 256 > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out/ | tail -12
