#!/bin/bash
# build an experiment variant of the library: tools/build_variant.sh <name> [-DFOO=1 ...]  -> gpurun_out/lib_<name>.so (use with TFR_LIB=...)
cd "$(dirname "$0")/.." && name=$1 && shift && mkdir -p gpurun_out/variants && \
python - "$name" "$@" <<'PY'
import sys, subprocess, os
import __graft_entry__ as g
name, extra = sys.argv[1], sys.argv[2:]
out = f"variants/lib_{name}.so"
subprocess.check_call([os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")] + g.NVCC_FLAGS + extra + ["-o", out, "spark-tfrecord_b200/csrc/api.cu"])
print("built", out)
PY
