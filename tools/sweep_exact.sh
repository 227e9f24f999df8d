#!/bin/bash
# usage: tools/sweep_exact.sh "<nvcc -D flags for exact_kernels.cu>" [bench args...]
# (run on the GPU box: rebuild only the streaming exact translation unit with the flags, then a short bench)
flags="$1"; shift
cp cgmres_cpp_b200/csrc/Makefile /tmp/Makefile.bak
touch cgmres_cpp_b200/csrc/exact_kernels.cu
make -j8 -C cgmres_cpp_b200/csrc EXACT_FMAD="-fmad=false $flags" > /dev/null 2>&1 || { echo "build failed: $flags"; exit 1; }
spill=$(grep -o "[0-9]* bytes spill stores" cgmres_cpp_b200/_build/exact_kernels.ptxas.log | sort -n | tail -1 | tr -d '\n')
timeout 300 python bench.py --mode exact --steps 30 --warmup 3 --no-cpu-baseline --no-other-modes "$@" 2>&1 | tail -1 > /tmp/sweep_line.json
python - "$flags" "$spill" <<'PY'
import json, sys
d = json.loads(open('/tmp/sweep_line.json').read())
print('%-44s max %-24s value %.4e  ms/step %.3f' % (sys.argv[1], sys.argv[2], d['value'], d['ms_per_step']))
PY
