set -x
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_r01b.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_r01b.log 2>&1
timeout 600 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --batch-nq 0 > gpurun_out/plain_r01b_scan.log 2>&1 || exit 1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:scan_small -s 3 -c 1 -o gpurun_out/prof_scan_small_q8 -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --batch-nq 0 > gpurun_out/ncu_scan_small.log 2>&1
echo done
