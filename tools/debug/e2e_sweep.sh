export BAMGPU_BENCH_CACHE=/dev/shm/bamgpu_cache
BAMGPU_DEBUG=2 timeout 300 python tools/debug/e2e_timeline.py 50000000 bodies > gpurun_out/timeline_bodies.txt 2>&1
tail -3 gpurun_out/timeline_bodies.txt
