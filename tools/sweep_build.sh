#!/bin/bash
# usage: tools/sweep_build.sh "<nvcc -D flags>" [bench args...]   (run on the GPU box: rebuild kernels, bench)
flags="$1"; shift
touch cgmres_cpp_b200/csrc/exact_kernels.cu cgmres_cpp_b200/csrc/onchip_exact_kernels.cu cgmres_cpp_b200/csrc/fast_kernels.cu
make -j8 -C cgmres_cpp_b200/csrc EXTRA="$flags" > /dev/null 2>&1 || { echo "build failed: $flags"; exit 1; }
spill=$(cat cgmres_cpp_b200/_build/*.ptxas.log | grep -o "[0-9]* bytes spill stores" | sort -n | tail -1 | tr -d '\n')
python bench.py --steps 40 --warmup 4 --no-cpu-baseline --no-other-modes "$@" 2>&1 | tail -1 > /tmp/sweep_line.json
python - "$flags" "$spill" <<'PY'
import json, sys
d = json.loads(open('/tmp/sweep_line.json').read())
print('%-44s max %-24s value %.4e  ms/step %.3f' % (sys.argv[1], sys.argv[2], d['value'], d['ms_per_step']))
PY
