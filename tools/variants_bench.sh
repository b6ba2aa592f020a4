#!/bin/bash
# tools/variants_bench.sh TAG:LIB ... -- quick_bench.sh for several builds of the library in one GPU call
# (tuning aid; LIB "-" = the in-tree library).
mkdir -p gpurun_out
for spec in "$@"; do
    tag=${spec%%:*}; lib=${spec#*:}
    if [ "$lib" = "-" ]; then lib=""; fi
    bash tools/quick_bench.sh $tag $lib
done 2>&1 | tee gpurun_out/variants_summary.txt
