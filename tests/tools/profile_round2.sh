# Round-2 profile (B200_PROFILING.md recipe): for each config the bench command runs plain first and, only if it exits 0,
# again under ncu for the launch list (per-launch device times: cold-cache and serialised, compare SHARES). AAADMM_NO_GRAPH=1:
# the loop body as plain stream launches (same kernels, same order) so that every kernel is its own launch for ncu.
mkdir -p gpurun_out
export AAADMM_NO_GRAPH=1
python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/r02_plain_cfg4.json 2> gpurun_out/r02_plain_cfg4.err && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1100 -c 220 --csv --log-file gpurun_out/r02_launches_cfg4.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/r02_ncu_cfg4.log 2>&1
echo "cfg4 ncu rc=$?"
python bench.py --config cfg3 --steps 1 --warmup 1 --no-cpu > gpurun_out/r02_plain_cfg3.json 2> gpurun_out/r02_plain_cfg3.err && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1200 -c 160 --csv --log-file gpurun_out/r02_launches_cfg3.csv python bench.py --config cfg3 --steps 1 --warmup 1 --no-cpu > gpurun_out/r02_ncu_cfg3.log 2>&1
echo "cfg3 ncu rc=$?"
python bench.py --config cfg1 --steps 1 --warmup 1 --no-cpu > gpurun_out/r02_plain_cfg1.json 2> gpurun_out/r02_plain_cfg1.err && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 3000 -c 200 --csv --log-file gpurun_out/r02_launches_cfg1.csv python bench.py --config cfg1 --steps 1 --warmup 1 --no-cpu > gpurun_out/r02_ncu_cfg1.log 2>&1
echo "cfg1 ncu rc=$?"
python bench.py --config cfg2 --steps 1 --warmup 1 --no-cpu > gpurun_out/r02_plain_cfg2.json 2> gpurun_out/r02_plain_cfg2.err && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1500 -c 160 --csv --log-file gpurun_out/r02_launches_cfg2.csv python bench.py --config cfg2 --steps 1 --warmup 1 --no-cpu > gpurun_out/r02_ncu_cfg2.log 2>&1
echo "cfg2 ncu rc=$?"
ls -la gpurun_out/r02_launches_* 
