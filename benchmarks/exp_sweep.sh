#!/usr/bin/env bash
# Round-2 sweep experiment: does thread-level parallelism help?  Same kernel at 32 / 16 / 8 dims per thread
# (8 / 16 / 12-16 warps per SM), same number of particle-dims.
set -u
mkdir -p gpurun_out
R7=gpurun_tmp/lib_r7/libaspire_b200.so
run() { # lib dims n moves tag
  ASPIRE_B200_LIB=$1 timeout 120 python benchmarks/micro.py sweep --iters 3 --only rosenbrock --dims $2 --n $3 --moves $4 2>&1 | grep '^{' | python -c "import sys,json; [print('$5', 'dims $2 n $3 moves $4', round(json.loads(l)['ms'],3), 'ms acc', round(json.loads(l)['acceptance'],5)) for l in sys.stdin]"
}
for lib in "" $R7; do
run "$lib" 32 20000000 10 "lib=$lib"
run "$lib" 16 40000000 10 "lib=$lib"
run "$lib" 8 80000000 10 "lib=$lib"
run "$lib" 4 160000000 10 "lib=$lib"
run "$lib" 32 20000000 20 "lib=$lib"
run "$lib" 16 40000000 20 "lib=$lib"
run "$lib" 8 80000000 20 "lib=$lib"
run "$lib" 4 160000000 20 "lib=$lib"
done
