set -x
mkdir -p gpurun_out/f2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ --csv --log-file gpurun_out/f2/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/f2/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"k_signbits|k_rowcount_blk|k_rowfill_cnt|k_cell_tris|k_scan_entries|k_cand_pos|k_seg_sort|k_unique|k_emit_faces" --launch-skip 22 -c 11 -f -o gpurun_out/f2/key1024 python tools/prof_kernels.py 1024 > gpurun_out/f2/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"k_signbits" --launch-skip 2 -c 1 -f -o gpurun_out/f2/signbits512 python tools/prof_mc.py 512 4 torus > gpurun_out/f2/ncu_sb512.log 2>&1
(echo "## memcheck"; timeout 1200 compute-sanitizer --tool memcheck python tools/sanitize_cases.py 2>&1 | tail -4; echo "## racecheck"; timeout 1200 compute-sanitizer --tool racecheck python tools/sanitize_cases.py 2>&1 | tail -4) > gpurun_out/f2/sanitizer.txt
timeout 300 python tools/bench_extra.py c5big > gpurun_out/f2/c5big.jsonl 2>/dev/null
ls -la gpurun_out/f2
