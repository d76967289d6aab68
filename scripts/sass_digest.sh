#!/bin/bash
# md5 of the device code (SASS of every kernel) inside libtesseract_b200.so, with the path-derived pieces (identifier lines,
# anonymous-namespace hashes) normalised: two builds with the same digest run the same kernels, whatever changed on the host.
LIB=${1:-$(dirname "$0")/../tesseract_decoder_b200/libtesseract_b200.so}
cuobjdump -sass "$LIB" | sed -E 's/_GLOBAL__N__[0-9a-f]{8}/_GLOBAL__N__X/g; /^identifier/d' | md5sum | cut -d' ' -f1
