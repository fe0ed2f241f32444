set -x
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r02F_pytest.log 2>&1; tail -2 gpurun_out/r02F_pytest.log
timeout 600 python bench.py > gpurun_out/r02F_bench_default.json 2> gpurun_out/r02F_bench_default.err; cut -c1-200 gpurun_out/r02F_bench_default.json
timeout 200 python bench.py --fpfh 1 --pairs 1024 --no-per-pair --no-cpu-baseline > gpurun_out/r02F_bench_fpfh.json 2> gpurun_out/r02F_bench_fpfh.err; cut -c1-200 gpurun_out/r02F_bench_fpfh.json
timeout 200 python bench.py --exact 0 --no-per-pair --no-cpu-baseline > gpurun_out/r02F_bench_treesums.json 2> gpurun_out/r02F_bench_treesums.err; cut -c1-200 gpurun_out/r02F_bench_treesums.json
timeout 100 python bench.py --steps 2 --warmup 1 --no-per-pair --no-cpu-baseline > gpurun_out/r02F_plain.log 2>&1 && timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02F_launches.csv python bench.py --steps 2 --warmup 1 --no-per-pair --no-cpu-baseline > gpurun_out/r02F_ncu_l.log 2>&1
timeout 100 python bench.py --pairs 1024 --steps 1 --warmup 1 --no-per-pair --no-cpu-baseline > gpurun_out/r02F_plain2.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_kernel -c 1 -o gpurun_out/r02F_search python bench.py --pairs 1024 --steps 1 --warmup 1 --no-per-pair --no-cpu-baseline > gpurun_out/r02F_ncu_s.log 2>&1
timeout 100 python scripts/one_register.py bunny300 > gpurun_out/r02F_bunny_plain.log 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:inner_bnb -s 20 -c 2 -o gpurun_out/r02F_bunny300_inner python scripts/one_register.py bunny300 > gpurun_out/r02F_ncu_b.log 2>&1
ls -la gpurun_out/r02F*
