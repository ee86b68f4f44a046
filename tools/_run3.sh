python -m pytest tests/test_gpu_euler_stage.py tests/test_gpu_halo_parity.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02_split_tests3.log
tools/order_sweep.sh "1 2 3 4 5 6 7 8" split3 > gpurun_out/r02_split_sweep3.log 2>&1
for v in mb4_4 fmb5 fmb6; do HDG_LIB_PATH=hopefoam_b200/variants/libhopedg_$v.so tools/order_sweep.sh "2 4" $v >> gpurun_out/r02_split_sweep3.log 2>&1; done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:euler -c 8 --csv --log-file gpurun_out/launches_split_r02b.csv python bench.py --order 4 --no-cpu --no-advection --steps 2 --warmup 1 --min-time 0.01 --e2e-steps 4 --e2e-serial > /dev/null 2>&1
HDG_LIB_PATH=hopefoam_b200/variants/libhopedg_fmb6.so ncu --metrics gpu__time_duration.sum --clock-control none -k regex:euler -c 8 --csv --log-file gpurun_out/launches_split_r02b_fmb6.csv python bench.py --order 4 --no-cpu --no-advection --steps 2 --warmup 1 --min-time 0.01 --e2e-steps 4 --e2e-serial > /dev/null 2>&1
cat gpurun_out/r02_split_tests3.log gpurun_out/r02_split_sweep3.log; tail -8 gpurun_out/launches_split_r02b.csv | cut -d, -f5,15; tail -8 gpurun_out/launches_split_r02b_fmb6.csv | cut -d, -f5,15
