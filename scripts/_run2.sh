out=gpurun_out
python bench.py --workload c3 --steps 200 --warmup 5 > $out/r2v_bench_c3.json 2> $out/r2v_bench_c3.err
python bench.py --workload c2 --steps 20 --warmup 3 > $out/r2v_bench_c2.json 2> $out/r2v_bench_c2.err
python scripts/optimizers_timing.py > $out/r2v_optimizers.txt 2>&1
cut -c1-400 $out/r2v_bench_c3.json; cut -c1-300 $out/r2v_bench_c2.json; tail -5 $out/r2v_optimizers.txt
