#!/bin/bash
# Round-end profile capture, run on the GPU box:   gpurun -- 'bash tools/gpu_capture.sh rNN'
# Brings back (gpurun_out/): profile_info.json, launches.csv, <tag>_*.ncu-rep, timelines, sanitizer logs.
# Summarise here with tools/ncu_summarise.py.
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
python tests/gpu_profile_render.py 3 3 > $out/profile_info.json 2> $out/profile_info.err
# launch list of one timed bench step (skip the warm-up forwards)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 170 -c 1500 --csv --log-file $out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-drift > $out/bench_under_ncu.log 2>&1
# full-set captures: second forward, middle (heaviest) 131072-ray chunk
cap() { # name regex skip [count] [driver...]
    local name=$1 re=$2 skip=$3 cnt=${4:-1}; shift 4
    timeout 400 ncu --set full --clock-control none --import-source on -k regex:$re -s $skip -c $cnt -f -o $out/${tag}_$name \
        "${@:-python tests/gpu_profile_render.py 3 0}" > $out/ncu_$name.log 2>&1
}
cap k_stage_q0 k_stage_q0 7 1 python tests/gpu_profile_render.py 3 0
cap k_stage_mid k_stage_mid 7 1 python tests/gpu_profile_render.py 3 0
cap k_nerf_mlp_coarse k_nerf_mlp 14 1 python tests/gpu_profile_render.py 3 0
cap k_nerf_mlp_fine k_nerf_mlp 15 1 python tests/gpu_profile_render.py 3 0
cap k_cconv k_cconv_tc 2 2 python tests/gpu_profile_trans.py 3           # second whole step: conv1 <96,64>, conv2 <64,64>
cap k_mlp_bwd k_mlp_bwd 2 2 python tests/gpu_profile_bwd.py              # second training step: dgrad, wgrad of the coarse net
# MLP tile timelines (tuning build: the same kernels + trace hooks)
export NF_B200_LIB=$PWD/neurofluid_b200/libnf_b200_tune.so
( echo "== k_nerf_mlp (one tile per CTA; tuning build only, NF_MLP_IMPL=1)"; NF_MLP_IMPL=1 timeout 120 python tests/gpu_mlp_trace.py;
  echo "== k_nerf_mlp2 (two tiles per CTA, production)"; timeout 120 python tests/gpu_mlp_trace2.py;
  echo "== sustained, back to back (clock / power sampled)"; REPS=40 timeout 200 python tests/gpu_mlp_power.py ) > $out/${tag}_mlp_timeline.txt 2>&1
unset NF_B200_LIB
# transition phases
timeout 200 python tests/gpu_trans_phases.py > $out/${tag}_trans_phases.txt 2>&1
# memcheck over the backward / transition / operator tests and the small render cases (every kernel of the library runs)
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_backward.py tests/test_gpu_transition.py -q -x \
    -k "not full_size and not end2end and not release" > $out/${tag}_sanitizer_memcheck.log 2>&1
tail -5 $out/${tag}_sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_render.py -q -x \
    -k "not full_size and not production" > $out/${tag}_sanitizer_memcheck_render.log 2>&1
tail -5 $out/${tag}_sanitizer_memcheck_render.log
ls -la $out | tail -30
