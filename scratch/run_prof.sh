set -x
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r01c.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1
python scratch/prof_kernels.py > gpurun_out/plain_prof.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:sed_broadband|gemm_3xtf32|dust_spectra|hist_kernel|knn_query|plan_kernel|expand_kernel|deposit_kernel|tile_reduce|select_hist|seg_reduce|item_scan' -c 44 -o gpurun_out/prof_r01c python scratch/prof_kernels.py > gpurun_out/ncu_prof.log 2>&1
echo finished
