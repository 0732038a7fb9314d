set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2f_tests.log 2>&1; tail -3 gpurun_out/r2f_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2f_smoke.log 2>&1; tail -2 gpurun_out/r2f_smoke.log
timeout 600 python bench.py > gpurun_out/r2f_bench_default_C4.json 2> gpurun_out/r2f_bench_default.err; tail -c 400 gpurun_out/r2f_bench_default_C4.json
timeout 300 python bench.py --workload C2 --no-cpu-baseline > gpurun_out/r2f_bench_C2.json 2> gpurun_out/r2f_bench_C2.err; tail -c 200 gpurun_out/r2f_bench_C2.json
timeout 400 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err; tail -c 300 gpurun_out/r2f_bench_ref.json
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2f_b.json 2> gpurun_out/r2f_b.err && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2f_ncu1.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_eval|k_asm_pose|k_pix|k_place|k_seg_sort" --launch-skip 14 --launch-count 7 -f -o gpurun_out/r2f_full_pass python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2f_ncu2.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_schur_tiles|k_ldlt_fused|k_solve_x2" --launch-skip 3 --launch-count 3 -f -o gpurun_out/r2f_full_solve python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2f_ncu3.log 2>&1
ls -la gpurun_out/ | tail -12
