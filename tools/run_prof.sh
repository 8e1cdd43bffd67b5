#!/bin/bash
# profiles of one round (run under gpurun on ONE GPU): plain run first, then the ncu launch list of the same command,
# then one full-set capture of the hot kernels; summaries are written by tools/launch_table.py / tools/ncu_table.py
R=${1:-r02}
set -x
python bench.py --steps 3 --warmup 3 --no-cpu --sections sed,spectra,imaging > gpurun_out/${R}_plain_bench.json 2> gpurun_out/${R}_plain_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --sections sed,spectra,imaging > gpurun_out/${R}_ncu_bench.log 2>&1
python tools/prof_kernels.py > gpurun_out/${R}_plain_prof.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on \
    -k 'regex:sed_broadband|gemm_3xtf32|dust_spectra|hist_kernel|knn_query|plan_kernel|expand_kernel|deposit_kernel|tile_reduce|select_hist|weights_kernel|knn_key|knn_gather|seg_reduce|recentre_chain' \
    -c 40 -o gpurun_out/${R}_prof python tools/prof_kernels.py > gpurun_out/${R}_ncu_prof.log 2>&1
ncu -i gpurun_out/${R}_prof.ncu-rep --page raw --csv > gpurun_out/${R}_prof_raw.csv 2>/dev/null
ls -la gpurun_out/ | tail -8
if [ $(stat -c %s gpurun_out/${R}_prof.ncu-rep) -gt 45000000 ]; then rm gpurun_out/${R}_prof.ncu-rep; fi
echo finished
