set -x
mkdir -p gpurun_out/f8
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > gpurun_out/f8/gputests.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f8/smoke.txt 2>&1
timeout 900 python bench.py > gpurun_out/f8/bench_n1.json 2> gpurun_out/f8/bench_n1.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/f8/bench_ref.json 2> gpurun_out/f8/bench_ref.err
for a in "512 torus mc" "1024 torus mc" "2048 csg mc" "512 csg dc"; do timeout 300 python tools/detail_timing.py $a >> gpurun_out/f8/detail.txt 2>&1; done
timeout 300 python tools/detail_sparse.py >> gpurun_out/f8/detail.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ --csv --log-file gpurun_out/f8/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sub-records > gpurun_out/f8/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"k_signbits|k_rowcount_blk|k_rowfill_cnt|k_cell_tris|k_scan_entries|k_cand_pos|k_seg_sort|k_unique|k_emit_faces" --launch-skip 22 -c 11 -f -o gpurun_out/f8/key1024 python tools/prof_kernels.py 1024 > gpurun_out/f8/ncu_full.log 2>&1
timeout 600 python tools/bench_extra.py c1 c4 c5 c5big > gpurun_out/f8/bench_extra.jsonl 2> gpurun_out/f8/bench_extra.err
cat gpurun_out/f8/gputests.txt gpurun_out/f8/smoke.txt
