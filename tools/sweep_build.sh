#!/bin/bash
# usage: tools/sweep_build.sh "<nvcc -D flags>" [bench args...]   (run on the GPU box: rebuild exact kernels, bench msd)
flags="$1"; shift
touch cgmres_cpp_b200/csrc/exact_kernels.cu
make -C cgmres_cpp_b200/csrc EXTRA="$flags" > /dev/null 2>&1 || { echo "build failed: $flags"; exit 1; }
spill=$(grep -A2 "control_kernelINS_21MassSpringDamperModelENS_25MassSpringDamperSimulatorELb0" cgmres_cpp_b200/_build/exact_kernels.ptxas.log | grep -o "[0-9]* bytes spill stores")
python bench.py --steps 40 --warmup 4 --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('%-44s %-24s value %.4e  ms/step %.3f' % ('$flags', '$spill', d['value'], d['ms_per_step']))"
