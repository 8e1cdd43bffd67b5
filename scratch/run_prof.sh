set -x
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r01f.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1
python scratch/prof_kernels.py > gpurun_out/plain_prof.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:sed_broadband|gemm_3xtf32|dust_spectra|hist_kernel|knn_query|plan_kernel|expand_kernel|deposit_kernel|tile_reduce|select_hist|select_compact|select_tail|weights_kernel|knn_key|knn_gather|seg_reduce|item_scan' -c 48 -o gpurun_out/prof_r01f python scratch/prof_kernels.py > gpurun_out/ncu_prof.log 2>&1
ncu -i gpurun_out/prof_r01f.ncu-rep --page raw --csv > gpurun_out/prof_r01f_raw.csv 2>/dev/null
ls -la gpurun_out/
if [ $(stat -c %s gpurun_out/prof_r01f.ncu-rep) -gt 45000000 ]; then rm gpurun_out/prof_r01f.ncu-rep; fi
echo finished
