#!/bin/bash
# Round-2 measurement artefacts from ONE build, on one B200 (run under gpurun from the repo root):
#   bash profiles/collect_r02.sh
# Everything lands in gpurun_out/; profiles/summarise_r02.py turns it into the committed profiles/r02_* files.
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/r02_pytest_gpu.log 2>&1; tail -2 $O/r02_pytest_gpu.log
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/r02_bench_reference_arm.json 2> $O/r02_bench_reference_arm.err
for w in c3 c1 c2 c5; do
  timeout 600 python bench.py --workload $w > $O/r02_bench_$w.json 2> $O/r02_bench_$w.err || echo "FAIL bench $w"
done
timeout 900 python bench.py --workload c4 > $O/r02_bench_c4.json 2> $O/r02_bench_c4.err || echo "FAIL bench c4"
timeout 300 python profiles/reg_compare.py c3 > $O/r02_reg_compare_c3.json 2> $O/r02_reg_compare_c3.err
# launch list of the bench command (cold-cache, serialised: shares, not absolutes).  The first call fills the keyframe
# cache (~3000 launches, once per session): the summary only counts what follows the last cache-fill kernel.
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --sustain-seconds 0 > $O/r02_plain_bench.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $O/r02_launches_bench_c3.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --sustain-seconds 0 > $O/r02_ncu_launches.log 2>&1
# full captures of the dominant kernels: registration, the bucket kernel of the surf map, one large radix-sort pass
timeout 300 python profiles/ncu_reg.py c3 3 > $O/r02_plain_reg.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:register_warm -s 1 -c 1 -o $O/r02_ncu_register -f \
    python profiles/ncu_reg.py c3 3 > $O/r02_ncu_register.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vgb_bucket_kernel -s 2 -c 1 -o $O/r02_ncu_bucket -f \
    python profiles/ncu_reg.py c3 3 > $O/r02_ncu_bucket.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:.*rs_onesweep_kernel<.int.512>.*" -s 8 -c 1 -o $O/r02_ncu_sort_new -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --sustain-seconds 0 > $O/r02_ncu_sort.log 2>&1
[ -s $O/r02_ncu_sort_new.ncu-rep ] && mv $O/r02_ncu_sort_new.ncu-rep $O/r02_ncu_sort.ncu-rep
ls -la $O | grep r02_ | head -40
