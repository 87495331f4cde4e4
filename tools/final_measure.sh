set -x
mkdir -p gpurun_out/f1
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/f1/gputests.txt
timeout 900 python bench.py > gpurun_out/f1/bench_n1.json 2> gpurun_out/f1/bench_n1.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/f1/bench_ref.json 2> gpurun_out/f1/bench_ref.err
for a in "512 torus mc" "1024 torus mc" "2048 csg mc" "512 csg dc"; do timeout 300 python tools/detail_timing.py $a >> gpurun_out/f1/detail.txt 2>&1; done
timeout 300 python tools/detail_sparse.py >> gpurun_out/f1/detail.txt 2>&1
timeout 900 python tools/bench_extra.py c1 c4 c5 c5big > gpurun_out/f1/bench_extra.jsonl 2> gpurun_out/f1/bench_extra.err
timeout 600 python tools/bench_sdf.py > gpurun_out/f1/bench_sdf.jsonl 2> gpurun_out/f1/bench_sdf.err
tail -3 gpurun_out/f1/gputests.txt
