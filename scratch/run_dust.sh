set -e
python -m pytest tests/test_gpu_sed.py -x -q -m gpu 2>&1 | tail -3
M=gpu__time_duration.sum,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed.sum
for h in 0 1; do
  SO_DUST_HYBRID=$h python scratch/bench_dust.py 2>&1 | tail -1
  SO_DUST_HYBRID=$h ncu --metrics $M --clock-control none -k regex:dust_spectra -c 1 --csv --log-file gpurun_out/dust_m$h.csv python scratch/bench_dust.py > gpurun_out/ncu_d$h.log 2>&1 || true
done
rm -f gpurun_out/dust_total_*.npy
