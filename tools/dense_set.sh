echo "G1 fwd (dgrad form) N=128"; python tools/one_kernel.py 1 128 1 8192 100 20 1 1 0
echo "G1 wgrad N=128"; python tools/one_kernel.py 2 128 1 8192 100 20 1 1 0
echo "D5 fwd N=256"; python tools/one_kernel.py 0 256 1 8192 1 20 1 1 0
echo "D5 dgrad N=256"; python tools/one_kernel.py 1 256 1 8192 1 20 1 1 0
echo "D5 wgrad N=256"; python tools/one_kernel.py 2 256 1 8192 1 20 1 1 0
