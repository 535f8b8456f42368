# round-2 experiment S: pageable caller buffers -- staging threads and chunk size
nproc; lscpu | grep -i "model name\|socket\|numa node(s)"
for t in 6 8 10; do for c in 1024 2048 4096 8192; do
J2K_STAGE_THREADS=$t J2K_STAGE_CHUNK_KB=$c timeout 200 python tools/pageable_probe.py 16 2>&1 | tail -1 | cut -c1-130
done; done
