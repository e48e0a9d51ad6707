# Developer tool: per-task trace of one LDL^T apply at cfg 4 (extra environment settings as arguments)
env "$@" AAADMM_LDLT_TRACE=1 timeout 300 python tests/tools/solve_only.py 148 37 37 3 > gpurun_out/trace_run.log 2>&1
gzip -f gpurun_out/ldlt_trace.csv
tail -2 gpurun_out/trace_run.log
