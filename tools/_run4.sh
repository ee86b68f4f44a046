python -m pytest tests/test_gpu_euler_stage.py tests/test_gpu_halo_parity.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02_split_tests4.log
tools/order_sweep.sh "1 2 3 4 5 6 7 8" split4 > gpurun_out/r02_split_sweep4.log 2>&1
for v in regbuf nopf noupdpf; do HDG_LIB_PATH=hopefoam_b200/variants/libhopedg_$v.so tools/order_sweep.sh "2 4 6" $v >> gpurun_out/r02_split_sweep4.log 2>&1; done
ncu --set full --clock-control none --import-source on -k regex:euler -s 8 -c 2 -o gpurun_out/prof_split_r02b -f python bench.py --order 4 --no-cpu --no-advection --steps 2 --warmup 1 --min-time 0.01 --e2e-steps 4 --e2e-serial > gpurun_out/ncu_split.log 2>&1
cat gpurun_out/r02_split_tests4.log gpurun_out/r02_split_sweep4.log
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02_gputests_c.log; cat gpurun_out/r02_gputests_c.log
