#!/bin/bash
# isolated timings of the C2 layer GEMMs (kind n h c o): D2,D3,D4 fprop/dgrad at N and 2N, wgrad at 2N
for k in 0 1; do for n in 128 256; do
  for g in "32 64 128" "16 128 256" "8 256 512"; do echo -n "kind=$k n=$n $g: "; python tools/one_kernel.py $k $n $g 20; done
done; done
for g in "32 64 128" "16 128 256" "8 256 512"; do echo -n "wgrad n=256 $g: "; python tools/one_kernel.py 2 256 $g 20; done
