#!/bin/bash
# Developer probe: the reference's test_multithread_stress over the GPU engine at several thread counts,
# with the engine's per-job trace counted (jobs per second, descriptors per call).
cd "$(dirname "$0")/.."
wd=$(mktemp -d); ln -s $PWD/power-gzip_b200/libnxz_gpu.so $wd/libnxz.so.1
for t in 1 8 64; do
  s=$(date +%s.%N)
  LD_LIBRARY_PATH=$wd:$PWD/power-gzip_b200 NX_GZIP_TYPE_SELECTOR=2 NX_GZIP_LOGFILE=$wd/nx.log NXGPU_TRACE=1 \
    oracle/_ref/reftests/test_multithread_stress $t 3 1 > $wd/out.$t 2> $wd/err.$t
  e=$(date +%s.%N)
  echo "threads $t: $(grep 'Total data' $wd/out.$t)  batches $(grep -c 'nxgpu batch' $wd/err.$t)  wall $(python3 -c "print(round($e-$s,2))") s"
  grep 'nxgpu batch' $wd/err.$t | awk '{d+=$3; b+=$5; u+=$(NF-1); n++} END {if (n) printf "   descriptors %d  (%.1f per batch)  source MB %.1f  mean batch %.0f us\n", d, d/n, b/1e6, u/n}'
  grep 'nxgpu batch' $wd/err.$t | awk '{print $8, $10}' | sort | uniq -c | sort -rn | head -6
  grep 'nxgpu batch' $wd/err.$t | sort -t, -k4 -n | tail -2
done
tail -3 $wd/nx.log 2>/dev/null
