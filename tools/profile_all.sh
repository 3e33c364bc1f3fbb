set -x
NCU="ncu --set full --clock-control none --import-source on"
python bench.py --steps 2 --warmup 3 --no-cpu --no-strict --no-configs > gpurun_out/prof_bench.json 2> gpurun_out/prof_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-strict --no-configs > gpurun_out/prof_ncu_bench.log 2>&1
$NCU -k regex:ghost_kernel -s 3 -c 1 -f -o gpurun_out/r2_ghost_v16 python tools/kernel_ab.py --cfg 2 --only pairs --reps 2 --precisions fp32 > gpurun_out/prof1.log 2>&1
$NCU -k regex:ghost_kernel -s 3 -c 1 -f -o gpurun_out/r2_ghost_strict python tools/kernel_ab.py --cfg 2 --only pairs --reps 2 --precisions strict > gpurun_out/prof2.log 2>&1
$NCU -k regex:prefix_kernel -s 3 -c 1 -f -o gpurun_out/r2_prefix python tools/kernel_ab.py --cfg 2 --only pairs --reps 2 --precisions fp32 > gpurun_out/prof3.log 2>&1
$NCU -k regex:family_kernel -s 1 -c 1 -f -o gpurun_out/r2_family_cfg3 python tools/kernel_ab.py --cfg 3 --only families --reps 1 --precisions fp32 > gpurun_out/prof4.log 2>&1
$NCU -k regex:tiles_kernel -s 5 -c 1 -f -o gpurun_out/r2_tiles_host python tools/sparse_loop.py 8 > gpurun_out/prof5.log 2>&1
$NCU -k regex:drain_kernel -s 6 -c 1 -f -o gpurun_out/r2_drain python tools/e2e_probe.py --steps 8 --slots 3 > gpurun_out/prof6.log 2>&1
$NCU -k regex:"star_pixels_kernel|zgemm_kernel|falloff" -c 4 -f -o gpurun_out/r2_starburst python -m pytest tests/test_starburst.py -m gpu -q -x -k "spectrum_cache" > gpurun_out/prof7.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -12
tail -2 gpurun_out/prof7.log
