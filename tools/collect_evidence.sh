#!/bin/bash
# Round evidence on ONE B200 (run through gpurun): tests, bench, launch lists, ncu --set full captures.
# Every ncu pass runs the same command that has just exited 0 without ncu.  Output: gpurun_out/ev_*
# (tools/make_profile_md.py <round> turns them into profiles/<round>_*).
O=gpurun_out
rm -f $O/ev_*
Q="--no-cpu-baseline --no-e2e --no-secondary --no-torch-gpu"
python -m pytest tests -m gpu -q > $O/ev_test.log 2>&1; tail -2 $O/ev_test.log
python bench.py --breakdown $O/ev_breakdown.txt > $O/ev_bench.log 2>&1; tail -1 $O/ev_bench.log | cut -c1-200
python bench.py --steps 2 --warmup 3 --batch 128 $Q > $O/ev_plain_b128.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/ev_launches.csv \
    python bench.py --steps 2 --warmup 3 --batch 128 $Q > $O/ev_ncu_list.log 2>&1
python bench.py --steps 2 --warmup 3 $Q > $O/ev_plain.log 2>&1 && {
ncu --set full --clock-control none --import-source on -k regex:conv3x3_sw -s 36 -c 5 -o $O/ev_prof_conv_sw -f \
    python bench.py --steps 2 --warmup 3 $Q > $O/ev_ncu_conv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_bres -s 3 -c 3 -o $O/ev_prof_gemm_bres -f \
    python bench.py --steps 2 --warmup 3 $Q > $O/ev_ncu_bres.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_out_mma -s 2 -c 1 -o $O/ev_prof_conv_out -f \
    python bench.py --steps 2 --warmup 3 $Q > $O/ev_ncu_convout.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_in_mma -s 2 -c 1 -o $O/ev_prof_conv_in -f \
    python bench.py --steps 2 --warmup 3 $Q > $O/ev_ncu_convin.log 2>&1
if [ "$EV_ALL" = 1 ]; then   # kernels unchanged since their last capture: re-captured only on request
ncu --set full --clock-control none --import-source on -k regex:ddpm_step -s 2 -c 1 -o $O/ev_prof_ddpm -f \
    python bench.py --steps 2 --warmup 3 $Q > $O/ev_ncu_ddpm.log 2>&1
fi
}
python tools/bench_configs.py > $O/ev_configs.json 2> $O/ev_configs.err
CDM_LIB=camels-diffusion-model_b200/libcdm_b200_probes.so python tools/gpu_probe.py hbm_rates waits waits_m4 perf_m3 perf_m4 perf_m4_c256 perf_m4_out0 perf_m4_pool wgrad wgrad_c256 > $O/ev_probe.log 2>&1
python tools/train_breakdown.py 256 > $O/ev_train_bd256.log 2>&1 && {
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/ev_train_launches.csv \
    python tools/train_breakdown.py 256 > $O/ev_ncu_train.log 2>&1
[ "$EV_ALL" = 1 ] && ncu --set full --clock-control none --import-source on -k regex:"bn_bwd_apply|chan_reduce_kernel|bn_apply_kernel" -s 40 -c 6 \
    -o $O/ev_prof_bn -f python tools/train_breakdown.py 256 > $O/ev_ncu_bn.log 2>&1
[ "$EV_ALL" = 1 ] && ncu --set full --clock-control none --import-source on -k regex:gemm_tn9_kernel -s 20 -c 1 -o $O/ev_prof_wgrad -f \
    python tools/train_breakdown.py 256 > $O/ev_ncu_wgrad.log 2>&1
}
ls $O | grep ev_
