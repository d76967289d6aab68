#!/bin/bash
# Host side of the library under AddressSanitizer + UndefinedBehaviorSanitizer: builds a second copy of
# libtesseract_b200.so in /tmp/tsr_san (host objects instrumented, CUDA objects as they are) and runs the host tests
# (DEM parser, ingestion, detector orders, circuit analyzer, sampler, fast-path certificate) against it.
# The near-tie DetCoordinate test is left out: it pins the production build's FMA flavour, which -O1 + UBSan changes.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SAN=/tmp/tsr_san
mkdir -p $SAN
FLAGS="-std=c++20 -O1 -g -fno-fast-math -ffp-contract=off -fPIC -Wall -Wextra -Wno-unused-parameter -pthread -fsanitize=address,undefined -fno-omit-frame-pointer"
make -s -j8 -C $ROOT/tesseract_decoder_b200/csrc OUT=$SAN BUILD=$SAN/_build CXXFLAGS="$FLAGS" \
  $SAN/_build/dem.o $SAN/_build/ingest.o $SAN/_build/detorder.o $SAN/_build/circuit.o $SAN/_build/fast.o \
  $SAN/_build/search.o $SAN/_build/search_fast.o $SAN/_build/capi.o $SAN/_build/group.o
ASAN_SO=$(gcc -print-file-name=libasan.so); UBSAN_SO=$(gcc -print-file-name=libubsan.so)
(cd $SAN/_build && nvcc -ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a -shared -o $SAN/libtesseract_b200.so \
  dem.o ingest.o detorder.o circuit.o fast.o search.o search_fast.o capi.o group.o -ldl -L$(dirname $ASAN_SO) -Xlinker -lasan -Xlinker -lubsan)
cat > $SAN/run.py <<PY
import sys
sys.path.insert(0, "$ROOT"); sys.path.insert(0, "$ROOT/tests")
from tesseract_decoder_b200 import capi
capi.LIB_PATH = "$SAN/libtesseract_b200.so"
import pytest
sys.exit(pytest.main(["$ROOT/tests/test_host.py", "-q", "-m", "not gpu", "-p", "no:cacheprovider", "-k", "not near_tied"]))
PY
ASAN_OPTIONS=detect_leaks=0:halt_on_error=0 UBSAN_OPTIONS=print_stacktrace=1 LD_PRELOAD="$ASAN_SO $UBSAN_SO" python $SAN/run.py 2>&1 | tee $SAN/out.log | tail -3
echo "sanitizer reports: $(grep -c 'runtime error\|AddressSanitizer' $SAN/out.log || true)"

# The pybind11 module and the CLI under the same sanitizers: a copy of the package with instrumented binaries.
PKG=$SAN/pkg/tesseract_decoder_b200
rm -rf $SAN/pkg && mkdir -p $PKG && cp $ROOT/tesseract_decoder_b200/*.py $PKG/ && cp $SAN/libtesseract_b200.so $PKG/
EXT=$(python -c "import sysconfig; print(sysconfig.get_config_var('EXT_SUFFIX'))")
SANFLAGS="-std=c++20 -O1 -g -fPIC -pthread -fsanitize=address,undefined -fno-omit-frame-pointer"
(cd $ROOT/tesseract_decoder_b200/csrc && /usr/bin/g++ $SANFLAGS -fvisibility=hidden -shared $(python -m pybind11 --includes) -I../../include pybind.cc \
   -o $PKG/_tesseract_b200$EXT -L$PKG -ltesseract_b200 -Wl,-rpath,'$ORIGIN' &&
 /usr/bin/g++ $SANFLAGS -I../../include main.cc -o $SAN/tesseract -L$SAN -ltesseract_b200 -Wl,-rpath,$SAN)
cat > $SAN/run_py.py <<PY
import sys
sys.path.insert(0, "$SAN/pkg"); sys.path.insert(1, "$ROOT"); sys.path.insert(2, "$ROOT/tests")
import tesseract_decoder_b200
assert tesseract_decoder_b200.__file__.startswith("$SAN/pkg")
import pytest
sys.exit(pytest.main(["$ROOT/tests/test_python_api.py", "$ROOT/tests/test_reference_python_tests.py", "-q", "-m", "not gpu", "-p", "no:cacheprovider"]))
PY
ASAN_OPTIONS=detect_leaks=0:halt_on_error=0 UBSAN_OPTIONS=print_stacktrace=1 LD_PRELOAD="$ASAN_SO $UBSAN_SO" python $SAN/run_py.py 2>&1 | tee $SAN/out_py.log | tail -3
echo "sanitizer reports (python API): $(grep -c 'runtime error\|AddressSanitizer' $SAN/out_py.log || true)"
echo "instrumented CLI for tests/test_cli.py-style runs: $SAN/tesseract"
