python -m cProfile -o gpurun_out/r2y_opt.prof scripts/optimizers_timing.py 250 lockstep > gpurun_out/r2y_opt_out.txt 2>&1
python -c "
import pstats
p=pstats.Stats('gpurun_out/r2y_opt.prof'); p.sort_stats('cumulative').print_stats(60)
" > gpurun_out/r2y_opt_profile.txt 2>&1
tail -3 gpurun_out/r2y_opt_out.txt | cut -c1-300
