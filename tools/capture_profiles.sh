#!/bin/bash
# (gpurun copies back at most 64 MiB: the .ncu-rep files are exported to CSV on the box and deleted)
# Round-2 profile capture (run on the GPU box through gpurun; ONE GPU).  Outputs under gpurun_out/prof/, summarised into profiles/ by
# tools/summarise_profiles.py.  Numbers printed by programs running under ncu are never bench values.
set -u
O=gpurun_out/prof; mkdir -p $O
NCU="ncu --set full --clock-control none --import-source on"
# 1. launch list of two whole steps (serialised, cold caches: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > $O/launches_bench.log 2>&1
# 2. full captures of the tensor-core kernels that lead the step, each through the production dispatch (tools/one_kernel.py)
cap() { name=$1; shift; kre=$1; shift; $NCU -k regex:$kre -s 1 -c 1 -o $O/$name python tools/one_kernel.py "$@" 3 > $O/$name.log 2>&1; ncu -i $O/$name.ncu-rep --page raw --csv > $O/$name.raw.csv 2>/dev/null; ncu -i $O/$name.ncu-rep --page details --csv > $O/$name.details.csv 2>/dev/null; rm -f $O/$name.ncu-rep; }
cap g4_forward_persist64 tc_conv_persistent 1 128 32 64 128
cap d2_dgrad_2n_persist64 tc_conv_persistent 1 256 32 64 128
cap d2_fprop_2n_persist128 tc_conv_persistent 0 256 32 64 128
cap d4_fprop_n_conv64 tc_conv_kernel 0 128 8 256 512
cap d3_fprop_2n_conv128 tc_conv_kernel 0 256 16 128 256
cap d2_wgrad_2n tc_wgrad 2 256 32 64 128
cap d3_wgrad_2n tc_wgrad 2 256 16 128 256
# 3. the HBM-bound kernels inside a real step
for k in bn_bwd_apply_acc_kernel bn_apply_acc_kernel updater_kernel reduce_multi_kernel; do
  $NCU -k regex:$k -s 4 -c 2 -o $O/step_$k python bench.py --steps 1 --warmup 3 --no-cpu --no-extra > $O/step_$k.log 2>&1
  ncu -i $O/step_$k.ncu-rep --page raw --csv > $O/step_$k.raw.csv 2>/dev/null; ncu -i $O/step_$k.ncu-rep --page details --csv > $O/step_$k.details.csv 2>/dev/null; rm -f $O/step_$k.ncu-rep
done
ls -la $O | head -40
