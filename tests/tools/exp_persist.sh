# Developer tool: A/B of the persistent sweep kernels (queue depth, trace)
AAADMM_SWEEP_QD=1 AAADMM_LDLT_TRACE=1 timeout 300 python tests/tools/solve_only.py 148 37 37 3 > gpurun_out/trace_run.log 2>&1
gzip -f gpurun_out/ldlt_trace.csv
tail -3 gpurun_out/trace_run.log
