# End-of-round evidence on one GPU: full GPU test suite, the bench lines of every config (with the reference as
# cpu_baseline), the reference arm, then the launch lists (tests/tools/profile_round2.sh).
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r02z_gputest.log; tail -1 gpurun_out/r02z_gputest.log
python bench.py > gpurun_out/r02z_bench_cfg4.json 2> gpurun_out/r02z_bench_cfg4.err; echo "cfg4 rc=$?"
for c in cfg1 cfg2 cfg3; do
  python bench.py --config $c > gpurun_out/r02z_bench_$c.json 2> gpurun_out/r02z_bench_$c.err; echo "$c rc=$?"
done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02z_reference_arm.json 2> gpurun_out/r02z_reference_arm.err; echo "reference rc=$?"
