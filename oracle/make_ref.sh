#!/usr/bin/env bash
# Ship the UNMODIFIED reference to the GPU box (test infrastructure, never product code).
#
#   oracle/make_ref.sh [/root/reference]
#
# 1. installs the reference package, as it is, into baseline/_ref (the one offline install the
#    build contract allows: pip --no-index --no-deps --target; the source tree is read-only, so
#    the wheel is built from a copy under /tmp);
# 2. copies the reference's own test module for the hot path (tests/test_string.py; tests/test_decoding.py for the step functions of the N-best producers), its
#    conftest.py and its pytest.ini (marker names) into oracle/_ref/tests/.
#
# Both directories are git-ignored (nothing of the reference enters this repository's history)
# but NOT gpurun-ignored, so they travel with the snapshot: tests/test_reference_suite.py runs
# the reference's tests through b200lev.install() on the B200, and bench.py --impl reference
# times the reference's own torch implementation.
set -euo pipefail
REF="${1:-/root/reference}"
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
if [ ! -d "$REF/src/pydrobert/torch" ]; then
    echo "make_ref: $REF is not the reference checkout; nothing to do" >&2
    exit 0
fi
PY="${PYTHON:-python}"
if [ ! -f "$ROOT/baseline/_ref/pydrobert/torch/_string.py" ] || [ "${FORCE:-0}" = 1 ]; then
    TMP="$(mktemp -d /tmp/b200lev_ref.XXXXXX)"
    cp -r "$REF" "$TMP/reference"
    rm -rf "$ROOT/baseline/_ref"
    "$PY" -m pip install --quiet --no-index --no-build-isolation --no-deps \
        --find-links /opt/wheelhouse --target "$ROOT/baseline/_ref" "$TMP/reference"
    rm -rf "$TMP"
fi
mkdir -p "$ROOT/oracle/_ref/tests"
cp "$REF/tests/test_string.py" "$REF/tests/test_decoding.py" "$REF/tests/conftest.py" "$ROOT/oracle/_ref/tests/"
# the reference registers its markers (cpu / gpu / trace / script / nojit) in pytest.ini
cp "$REF/pytest.ini" "$ROOT/oracle/_ref/tests/pytest.ini"
echo "make_ref: baseline/_ref (package) and oracle/_ref/tests (test_string.py, conftest.py) ready"
