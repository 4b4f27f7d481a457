set -x
python profiles/bench_planar_layers.py > gpurun_out/planar_layers_base.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv2d_tc -o /tmp/pl python profiles/bench_planar_layers.py --only 2,12,0,3 --launches 2 > gpurun_out/ncu_pl.log 2>&1
ncu -i /tmp/pl.ncu-rep --page raw --csv > gpurun_out/pl_raw.csv 2>/dev/null
for id in 1 3 5; do ncu -i /tmp/pl.ncu-rep --page source --csv --launch-skip $id --launch-count 1 > gpurun_out/pl_src_$id.csv 2>/dev/null; done
ls -la gpurun_out/ | tail -8
cat gpurun_out/planar_layers_base.txt
