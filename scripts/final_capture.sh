set -x
timeout 150 python bench.py > gpurun_out/r2P_c4_n1.log 2> gpurun_out/r2P_c4_n1.err; echo bench rc $?
timeout 100 python bench.py --steps 2 --warmup 3 --no-graph --no-e2e --no-cpu-baseline > gpurun_out/r2P_plain.log 2>&1; echo plain rc $?
timeout 110 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2P_launches.csv python bench.py --steps 2 --warmup 3 --no-graph --no-e2e --no-cpu-baseline > gpurun_out/r2P_ncu1.log 2>&1; echo ncu1 rc $?
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"gemm_nt_kernel|gemm_nt_ares_kernel|gemm_wgrad_kernel" -s 18 -c 6 -o /tmp/r2P_full python bench.py --steps 2 --warmup 3 --no-graph --no-e2e --no-cpu-baseline > gpurun_out/r2P_ncu2.log 2>&1; echo ncu2 rc $?
ncu -i /tmp/r2P_full.ncu-rep --page raw --csv > gpurun_out/r2P_ncu_full_raw.csv 2>/dev/null; ls -la gpurun_out/
