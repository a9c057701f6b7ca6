#!/usr/bin/env bash
# Stage the UNMODIFIED reference (esa/torchquad, /root/reference) under the git-ignored baseline/_ref/ so that it
# travels to the GPU box with the repo snapshot (git-ignored files do travel; /root/reference does not exist there).
#   baseline/_ref/torchquad/   `pip install --no-deps --target` of the reference (its dependency autoray is absent
#                              from the offline wheelhouse: oracle/autoray_standin provides it at import time)
#   baseline/_ref/tests/       the reference's own test-suite, run against torchquad_b200 by
#                              tests/test_gpu_reference_suite.py and against the reference itself by run_ref_tests.sh
# Nothing under baseline/_ref/ is ever committed (GPL sources; .gitignore lists it).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${1:-/root/reference}"
if [ ! -d "$REF/torchquad" ]; then
    echo "stage_ref: $REF/torchquad not found (on the GPU box the staged copy travels with the snapshot)" >&2
    exit 0
fi
if [ -f "$HERE/_ref/.staged" ] && diff -rq -x __pycache__ "$REF/torchquad" "$HERE/_ref/torchquad" >/dev/null 2>&1 \
   && diff -rq -x __pycache__ "$REF/tests" "$HERE/_ref/tests" >/dev/null 2>&1; then
    echo "stage_ref: baseline/_ref is up to date"
    exit 0
fi
rm -rf "$HERE/_ref"
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
cp -r "$REF" "$TMP/src"   # /root/reference is read-only and the build writes into the source tree
if ! python -m pip install --quiet --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
        --target "$HERE/_ref" "$TMP/src" >"$TMP/pip.log" 2>&1; then
    echo "stage_ref: pip install failed, copying the package directory instead" >&2
    cat "$TMP/pip.log" >&2
    mkdir -p "$HERE/_ref"
    cp -r "$REF/torchquad" "$HERE/_ref/torchquad"
fi
cp -r "$REF/tests" "$HERE/_ref/tests"
find "$HERE/_ref" -name __pycache__ -type d -prune -exec rm -rf {} +
diff -rq "$REF/torchquad" "$HERE/_ref/torchquad"   # must be byte-identical: the baseline is the UNMODIFIED reference
date -u +%FT%TZ > "$HERE/_ref/.staged"
echo "stage_ref: staged $(find "$HERE/_ref/torchquad" -name '*.py' | wc -l) package files and $(ls "$HERE/_ref/tests" | wc -l) test files"
