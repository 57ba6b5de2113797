#!/bin/bash
# Round-2 evidence run (one B200): bench lines, ncu launch list, ncu full capture of the edge kernel, k sweep of
# configs[4], the configs[2] job, compute-sanitizer on a small shape.  Every command has its own timeout.
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 170 python bench.py --steps 20 --warmup 5 > $O/r02_bench_default.json 2> $O/r02_bench_default.err
timeout 120 python bench.py --impl reference --steps 5 --warmup 1 > $O/r02_bench_reference.json 2> $O/r02_bench_reference.err
MPN_BENCH_NO_SAMPLER=1 timeout 170 ncu --metrics gpu__time_duration.sum --clock-control none -s 260 -c 140 --csv --log-file $O/r02_launches.csv \
  python bench.py --steps 2 --warmup 2 --no-cpu-baseline > $O/r02_launches.log 2>&1
MPN_BENCH_NO_SAMPLER=1 timeout 170 ncu --set full --clock-control none --import-source on -k regex:"mp_edge_tc3" -s 14 -c 1 -o $O/r02_edge -f \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/r02_edge_ncu.log 2>&1
MPN_BENCH_NO_SAMPLER=1 timeout 170 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed \
  --clock-control none -k regex:"gram_blocks2|node_encoder_tc|cand_select|row_threshold|node_tc2|mp_edge_tc3|avgpool|row_select|exact_rows|pair_dist|mask_pairs|bit_transpose" -s 60 -c 60 --csv --log-file $O/r02_kernel_metrics.csv \
  python bench.py --steps 2 --warmup 2 --no-cpu-baseline > $O/r02_kernel_metrics.log 2>&1
for k in 25 50 100 150; do
  timeout 100 python bench.py --steps 10 --warmup 3 --dets 300 --k $k --graphs 4 --pooled --no-cpu-baseline > $O/r02_config5_k$k.json 2> $O/r02_config5_k$k.err
done
timeout 200 python bench.py --steps 5 --warmup 2 --job-graphs 512 --pooled --no-cpu-baseline > $O/r02_config3_job512.json 2> $O/r02_config3_job512.err
timeout 200 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_sanitizer_memcheck.log 2>&1
timeout 200 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_sanitizer_racecheck.log 2>&1
tail -c 400 $O/r02_bench_default.json; echo; tail -2 $O/r02_sanitizer_memcheck.log; tail -2 $O/r02_sanitizer_racecheck.log; ls -la $O | grep r02_
