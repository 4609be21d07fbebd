export BAMGPU_BENCH_CACHE=/dev/shm/bamgpu_cache
for mode in bodies cols; do
for j in 4 8 12; do for mib in 128 256 512; do
echo "== mode $mode jobs $j batch $mib"; BAMGPU_JOBS=$j BAMGPU_BATCH_MIB=$mib timeout 300 python tools/debug/e2e_timeline.py 50000000 $mode 2>&1 | tail -3
done; done; done
BAMGPU_DEBUG=2 timeout 300 python tools/debug/e2e_timeline.py 50000000 bodies 2>&1 | tail -150 > gpurun_out/timeline_bodies.txt
