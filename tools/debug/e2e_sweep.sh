export BAMGPU_BENCH_CACHE=/dev/shm/bamgpu_cache
for mode in bodies cols; do
for j in 8; do for mib in 256; do
echo "== mode $mode jobs $j batch default"; BAMGPU_JOBS=$j timeout 300 python tools/debug/e2e_timeline.py 50000000 $mode 2>&1 | tail -3
done; done; done
