#!/bin/bash
# build a variant of the library with extra -D flags into variants/<name>/ (development A/B tests):
#   tools/build_variant.sh NAME [-DLZB_V6_DEV] [-D...]     only lanczos_v6.cu is recompiled with the flags
name=$1; shift
mkdir -p variants/$name
python - "$name" "$@" <<'PY'
import sys
from lanczos_hls_b200 import build
name, extra = sys.argv[1], sys.argv[2:]
print(build.build(extra=extra, out="variants/%s/liblanczos_b200.so" % name, only=["lanczos_v6.cu"], verbose="-v" in extra))
PY
