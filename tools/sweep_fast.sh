#!/bin/bash
# usage: tools/sweep_fast.sh "<nvcc -D flags for fast_kernels.cu>" [bench args...]
# (run on the GPU box: rebuild only the fast-mode translation unit with the flags, then a short bench)
flags="$1"; shift
touch cgmres_cpp_b200/csrc/fast_kernels.cu
make -j8 -C cgmres_cpp_b200/csrc FAST_EXTRA="$flags" > /dev/null 2>&1 || { echo "build failed: $flags"; exit 1; }
spill=$(grep -o "[0-9]* bytes spill stores" cgmres_cpp_b200/_build/fast_kernels.ptxas.log | sort -n | tail -1 | tr -d '\n')
timeout 300 python bench.py --steps 40 --warmup 4 --no-cpu-baseline --no-other-modes "$@" 2>&1 | tail -1 > /tmp/sweep_line.json
python - "$flags" "$spill" <<'PY'
import json, sys
d = json.loads(open('/tmp/sweep_line.json').read())
print('%-44s max %-24s value %.4e  ms/step %.3f' % (sys.argv[1], sys.argv[2], d['value'], d['ms_per_step']))
PY
