# End-of-round `ncu --set full` of the two sweep kernels at cfg 4 (plain run first, then ncu; raw page exported on the box)
mkdir -p gpurun_out
export AAADMM_NO_GRAPH=1
python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/r02_fullz_plain_cfg4.json 2> gpurun_out/r02_fullz_plain_cfg4.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_fwd_front|k_bwd_front" -s 20 -c 4 -f -o gpurun_out/r02_fullz_cfg4 python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/r02_fullz_ncu_cfg4.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/r02_fullz_cfg4.ncu-rep --page raw --csv > gpurun_out/r02_fullz_cfg4_raw.csv 2>/dev/null; rm -f gpurun_out/r02_fullz_cfg4.ncu-rep
ls -la gpurun_out/r02_fullz_*
