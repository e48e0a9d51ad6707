# Developer tool: ncu --set full on selected sweep kernels of one LDLT apply (raw CSV to gpurun_out/)
set -e
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:front -s ${1:-38} -c ${2:-4} -o /tmp/ldlt_full python tests/tools/solve_only.py 148 37 37 3 > gpurun_out/ncu_full.log 2>&1
ncu -i /tmp/ldlt_full.ncu-rep --page raw --csv > gpurun_out/ldlt_full_raw.csv
ls -la /tmp/ldlt_full.ncu-rep gpurun_out/ldlt_full_raw.csv
