# the committed Performer profiles of one code state, in one GPU call (outputs under gpurun_out/):
#   launch list of one step, `--set full` of the nine NT-GEMM launches of a depth-1 step, `--set full` of the fused FAVOR+ backward kernels
set -x
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_performer.csv \
    python tools/step_profile.py performer bf16 > gpurun_out/prof1.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tc_gemm_nt -c 9 -f -o gpurun_out/r2_gemm \
    python tools/step_profile.py performer bf16 --depth 1 > gpurun_out/prof2.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tc_dqk_fb -c 2 -f -o gpurun_out/r2_dqk_fb \
    python tools/step_profile.py performer bf16 --depth 1 > gpurun_out/prof3.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_performer.csv
