set -x
timeout 900 python bench.py > gpurun_out/bench_final_n1.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:scan_kernel -s 3 -c 1 -o gpurun_out/prof_scan_q8_final -f python bench.py --steps 1 --warmup 3 --nq 8 --no-e2e --no-cpu-baseline --batch-nq 0 > gpurun_out/ncu_scan.log 2>&1
echo done
