# Last evidence pass of round 2 on one GPU (after the final kernel / schedule changes): GPU tests, smoke, bench lines of
# every config, then the ncu launch lists of cfg 4 and cfg 3 (each bench command plain first, then under ncu).
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r02f_gputest.log; tail -1 gpurun_out/r02f_gputest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r02f_bench_cfg4.json 2> gpurun_out/r02f_bench_cfg4.err; echo "cfg4 rc=$?"
for c in cfg1 cfg2 cfg3; do python bench.py --config $c > gpurun_out/r02f_bench_$c.json 2> gpurun_out/r02f_bench_$c.err; echo "$c rc=$?"; done
timeout 600 python tests/tools/cloth_bench.py 1000 0 2>/dev/null > gpurun_out/r02f_cloth_2M_phases.json
export AAADMM_NO_GRAPH=1
python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/r02_plain_cfg4.json 2> gpurun_out/r02_plain_cfg4.err && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1100 -c 220 --csv --log-file gpurun_out/r02_launches_cfg4.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/r02_ncu_cfg4.log 2>&1
echo "cfg4 ncu rc=$?"
python bench.py --config cfg3 --steps 1 --warmup 1 --no-cpu > gpurun_out/r02_plain_cfg3.json 2> gpurun_out/r02_plain_cfg3.err && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1200 -c 160 --csv --log-file gpurun_out/r02_launches_cfg3.csv python bench.py --config cfg3 --steps 1 --warmup 1 --no-cpu > gpurun_out/r02_ncu_cfg3.log 2>&1
echo "cfg3 ncu rc=$?"
