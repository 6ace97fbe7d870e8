python scripts/variant_bench.py c3 > gpurun_out/r2v_vb.txt 2>&1
python scripts/variant_bench.py c3 122 >> gpurun_out/r2v_vb.txt 2>&1
python scripts/variant_bench.py c2 >> gpurun_out/r2v_vb.txt 2>&1
python scripts/variant_bench.py c4 >> gpurun_out/r2v_vb.txt 2>&1
python scripts/variant_bench.py c5 >> gpurun_out/r2v_vb.txt 2>&1
cat gpurun_out/r2v_vb.txt
