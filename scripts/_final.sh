out=gpurun_out; t=r2z
python -m pytest tests -m gpu -q > $out/${t}_pytest.log 2>&1; tail -2 $out/${t}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/${t}_smoke.log 2>&1; tail -1 $out/${t}_smoke.log
python bench.py > $out/${t}_bench_c4.json 2> $out/${t}_bench_c4.err
python bench.py --impl reference > $out/${t}_bench_ref.json 2> $out/${t}_bench_ref.err
python bench.py --workload c2 --steps 20 --warmup 3 > $out/${t}_bench_c2.json 2> $out/${t}_bench_c2.err
python bench.py --workload c3 --steps 200 --warmup 5 > $out/${t}_bench_c3.json 2> $out/${t}_bench_c3.err
python bench.py --workload c5 --steps 5 --warmup 3 > $out/${t}_bench_c5.json 2> $out/${t}_bench_c5.err
for w in c4 ref c2 c3 c5; do cut -c1-220 $out/${t}_bench_$w.json; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${t}_c4_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > $out/${t}_ncu_launches.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${t}_smoke_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > $out/${t}_ncu_smoke.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ttvb_main -s 2 -c 1 -f -o $out/${t}_c4 python scripts/variant_bench.py c4 > $out/${t}_ncu_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ttvb_main -s 2 -c 1 -f -o $out/${t}_c3 python scripts/variant_bench.py c3 > $out/${t}_ncu_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ttvb_main -s 2 -c 1 -f -o $out/${t}_c2 python scripts/variant_bench.py c2 > $out/${t}_ncu_c2.log 2>&1
ls -la $out | tail -20
