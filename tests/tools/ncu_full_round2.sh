# (the .ncu-rep files exceed what gpurun brings back: the raw page is exported as CSV on the box)
# Round-2 `ncu --set full` captures (B200_PROFILING.md recipe): the bench command runs plain first and, only if it exits 0,
# again under ncu. AAADMM_NO_GRAPH=1: the loop body as plain stream launches so that every kernel is its own launch.
mkdir -p gpurun_out
export AAADMM_NO_GRAPH=1
python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/r02_full_plain_cfg4.json 2> gpurun_out/r02_full_plain_cfg4.err && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_update_z_hard|k_fwd_front|k_bwd_front|k_aa_pass1|k_aa_pass2|k_update_u_hard|k_rhs_gather" -s 70 -c 14 -f -o gpurun_out/r02_full_cfg4 python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/r02_full_ncu_cfg4.log 2>&1
echo "cfg4 ncu rc=$?"
ncu -i gpurun_out/r02_full_cfg4.ncu-rep --page raw --csv > gpurun_out/r02_full_cfg4_raw.csv 2>/dev/null; rm -f gpurun_out/r02_full_cfg4.ncu-rep
python bench.py --config cfg3 --steps 1 --warmup 1 --no-cpu > gpurun_out/r02_full_plain_cfg3.json 2> gpurun_out/r02_full_plain_cfg3.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_geo_local|k_geo_soft|k_geo_rhs|k_geo_u_resid|k_fwd_front|k_bwd_front" -s 60 -c 12 -f -o gpurun_out/r02_full_cfg3 python bench.py --config cfg3 --steps 1 --warmup 1 --no-cpu > gpurun_out/r02_full_ncu_cfg3.log 2>&1
echo "cfg3 ncu rc=$?"
ncu -i gpurun_out/r02_full_cfg3.ncu-rep --page raw --csv > gpurun_out/r02_full_cfg3_raw.csv 2>/dev/null; rm -f gpurun_out/r02_full_cfg3.ncu-rep
python bench.py --config cfg1 --steps 1 --warmup 1 --no-cpu > gpurun_out/r02_full_plain_cfg1.json 2> gpurun_out/r02_full_plain_cfg1.err && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_hyper|k_grad_u_xzu|k_update_z_plain" -s 30 -c 6 -f -o gpurun_out/r02_full_cfg1 python bench.py --config cfg1 --steps 1 --warmup 1 --no-cpu > gpurun_out/r02_full_ncu_cfg1.log 2>&1
echo "cfg1 ncu rc=$?"
ncu -i gpurun_out/r02_full_cfg1.ncu-rep --page raw --csv > gpurun_out/r02_full_cfg1_raw.csv 2>/dev/null; rm -f gpurun_out/r02_full_cfg1.ncu-rep
ls -la gpurun_out/r02_full_*
