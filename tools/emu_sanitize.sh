#!/bin/bash
# tools/emu_sanitize.sh [address|thread] -- the kernel logic under a sanitizer on the CPU.
# compute-sanitizer is closed on this GPU pool, so the memcheck / racecheck evidence comes from the SIMT
# emulator build of the very same kernel sources (tests/emu): every CUDA thread is an OS thread there,
# shared memory and "device" memory are ordinary host memory, so AddressSanitizer sees every out-of-bounds
# or use-after-free access of the kernels and ThreadSanitizer every unsynchronised conflicting access
# between lanes.  Runs the emulator test-suite; the log goes to stdout.
KIND=${1:-address}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=/tmp/tess_emu_$KIND
mkdir -p $OUT
if [ "$KIND" = address ]; then FLAGS="-fsanitize=address,undefined -fno-sanitize-recover=undefined"; RT=$(g++ -print-file-name=libasan.so)
else FLAGS="-fsanitize=thread"; RT=$(g++ -print-file-name=libtsan.so); fi
g++ -std=c++20 -O1 -g -pthread $FLAGS -fno-omit-frame-pointer -DTESS_EMU -include $ROOT/tests/emu/cuda_emu.h -x c++ \
    $ROOT/tess_b200/csrc/tess_b200.cu -shared -fPIC -o $OUT/libtess_emu.so || exit 1
export TESS_EMU_LIB=$OUT/libtess_emu.so
export ASAN_OPTIONS=detect_leaks=0:abort_on_error=0:halt_on_error=1 TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0"
# -s: sanitizer reports go to stderr while a test runs; pytest would swallow them otherwise
LD_PRELOAD=$RT python -m pytest $ROOT/tests/test_emu_kernels.py -x -q -s -p no:cacheprovider "${@:2}" > $OUT/log.txt 2>&1
echo "== $KIND sanitizer, emulator build of tess_b200/csrc (g++ $FLAGS), tests/test_emu_kernels.py ${@:2}"
grep -E "passed|failed|error" $OUT/log.txt | tail -3
echo "== reports: $(grep -c -E 'WARNING: ThreadSanitizer|ERROR: AddressSanitizer|runtime error' $OUT/log.txt)"
# one line per distinct report site (first frame inside the kernel sources)
grep -A3 -E "WARNING: ThreadSanitizer|ERROR: AddressSanitizer" $OUT/log.txt | grep -E "#0 " | sed -E 's/\(libtess_emu.*//; s/^ +#0 //' | sort | uniq -c | sort -rn | head -20
grep -E "runtime error" $OUT/log.txt | sort | uniq -c | head
