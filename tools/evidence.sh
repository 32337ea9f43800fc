#!/bin/bash
# Round evidence on one B200 (run through gpurun): bench line, ncu launch list of the same command, `ncu --set full`
# captures of the dominant kernels of both stages in both modes.  Usage: bash tools/evidence.sh <tag>   (writes gpurun_out/<tag>_*)
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
timeout 900 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
tail -c 600 $out/${tag}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_ncu_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-fast --no-parity-gate --no-cpu-baseline --no-gpu-baseline > $out/${tag}_ncu_bench.log 2>&1
NCU="ncu --set full --metrics sm__inst_executed_pipe_tensor.sum --clock-control none --import-source on -f"
# stage 1: tensor-memory RIC kernel; its launches 0-16 of a forward are conv0-2 and the 14 trunk convolutions, 17 / 18 = upconv2 / upconv1
timeout 300 $NCU -k regex:conv_ric_tm_kernel -s 17 -c 2 -o $out/${tag}_s1_upconv2_upconv1_fp16x3 python tools/ncu_target.py 1 fp16x3 > $out/${tag}_ncu_a.log 2>&1
timeout 300 $NCU -k regex:conv_ric_tm_kernel -s 17 -c 2 -o $out/${tag}_s1_upconv2_upconv1_fp16 python tools/ncu_target.py 1 fp16 > $out/${tag}_ncu_b.log 2>&1
# stage 2: persistent halo kernel, launch 22 = conv_11 (7x7, 166 -> 64) after 14 trunk + 8 sub-pixel launches
timeout 300 $NCU -k regex:conv_halo_persist_kernel -s 22 -c 1 -o $out/${tag}_s2_conv11_fp16x3 python tools/ncu_target.py 2 fp16x3 > $out/${tag}_ncu_c.log 2>&1
timeout 300 $NCU -k regex:conv_halo_persist_kernel -s 22 -c 1 -o $out/${tag}_s2_conv11_fp16 python tools/ncu_target.py 2 fp16 > $out/${tag}_ncu_d.log 2>&1
ls -la $out/${tag}_*
