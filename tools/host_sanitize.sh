#!/bin/bash
# Host-side sanitizer pass, no GPU needed: the translation units that hold the library's host logic (C ABI argument handling, transcript,
# vk hash, permutation assembly, XorShift draw, point sums, tree / window layouts) are rebuilt with AddressSanitizer + UBSan, linked with the
# regular objects of the device-heavy units, and exercised by (1) the device-free half of examples/simple_example.cpp and (2) the CPU ABI tests.
# compute-sanitizer is closed on the GPU pool (profiles/r2_compute_sanitizer_closed.log); this covers the host half of the same concern.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=${1:-/tmp/h2a_asan}
B=$ROOT/halo2-aggregation_b200/build
mkdir -p "$OUT"
python "$ROOT/halo2-aggregation_b200/_build.py" >/dev/null
for f in glue keygen abi misc; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O1 -g -std=c++17 --expt-relaxed-constexpr \
    -Xcompiler -fPIC,-fsanitize=address,-fsanitize=undefined,-fno-omit-frame-pointer -c "$ROOT/halo2-aggregation_b200/csrc/$f.cu" -o "$OUT/$f.o" &
done
wait
/usr/local/cuda/bin/nvcc -shared -o "$OUT/libh2agg.so" "$OUT"/{glue,keygen,abi,misc}.o "$B"/{msm,ntt,plonk_verify,plonk_prove,params,comm,mulvar}.o \
  -gencode arch=compute_100a,code=sm_100a -lcudart -ldl -Xcompiler -fsanitize=address,-fsanitize=undefined
g++ -std=c++17 -g -fsanitize=address,undefined -I"$ROOT/include" "$ROOT/examples/simple_example.cpp" -L"$OUT" -lh2agg -Wl,-rpath,"$OUT" -o "$OUT/simple_example"
echo "== examples/simple_example.cpp --host-only (ASan + UBSan + LeakSanitizer)"
ASAN_OPTIONS=detect_leaks=1 "$OUT/simple_example" --host-only | tail -2
echo "== tests/test_abi_cpu.py against the sanitized library"
cd "$ROOT"
ASAN_OPTIONS=detect_leaks=0:verify_asan_link_order=0 LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" H2A_LIB="$OUT/libh2agg.so" \
  python -m pytest tests/test_abi_cpu.py -q -p no:cacheprovider -k "not compiles and not cpp and not rust and not generated" 2>&1 | tail -2
