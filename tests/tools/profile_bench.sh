# Round profile: (1) bench without ncu, (2) ncu launch list of the same command, (3) ncu --set full of the sweep kernels.
set -e
mkdir -p gpurun_out
AAADMM_NO_GRAPH=1 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/prof_bench_plain.json 2> gpurun_out/prof_bench_plain.err
AAADMM_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1100 -c 200 --csv --log-file gpurun_out/r01b_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/prof_ncu1.log 2>&1
AAADMM_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:front -s 20 -c 2 -o /tmp/r01b_full python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/prof_ncu2.log 2>&1
ncu -i /tmp/r01b_full.ncu-rep --page raw --csv > gpurun_out/r01b_full_ldlt_raw.csv
ls -la gpurun_out/
