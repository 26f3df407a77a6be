set -x
python bench.py --steps 2 --warmup 3 > gpurun_out/r02_bench_mid.json 2> gpurun_out/r02_bench_mid.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_mid.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_ncu_launches_mid.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name regex:'k_gram_warp|k_chol_solve|k_basis|k_interp_hw|k_synth' --launch-skip 12 --launch-count 12 -o gpurun_out/prof_r02_mid python tools/kernel_times.py 20 1 > gpurun_out/r02_ncu_full_mid.log 2>&1
ls -la gpurun_out/prof_r02_mid.ncu-rep
