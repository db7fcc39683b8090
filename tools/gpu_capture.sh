#!/bin/bash
# Round-end profile capture, run on the GPU box:   gpurun -- 'bash tools/gpu_capture.sh rNN'
# Brings back (gpurun_out/): profile_info.json, launches.csv, <tag>_*.ncu-rep.  Summarise here with tools/ncu_summarise.py.
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
python tests/gpu_profile_render.py 3 3 > $out/profile_info.json 2> $out/profile_info.err
# launch list of one timed bench step (skip the warm-up forwards: 5 x ~34 launches)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 170 -c 1500 --csv --log-file $out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/bench_under_ncu.log 2>&1
# full-set captures: second forward, middle (heaviest) 131072-ray chunk
cap() { # name regex skip [count]
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c ${4:-1} -f -o $out/${tag}_$1 \
        python tests/gpu_profile_render.py 3 3 > $out/ncu_$1.log 2>&1
}
cap k_stage_q0 k_stage_q0 7
cap k_stage_mid k_stage_mid 7
cap k_nerf_mlp_coarse k_nerf_mlp 14
cap k_nerf_mlp_fine k_nerf_mlp 15
cap k_cconv k_cconv_tc 3 3          # second transition step: conv1 <96,64>, conv2 <64,64>, conv3 <64,16>
ls -la $out
