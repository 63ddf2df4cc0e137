mkdir -p gpurun_out/r3u
timeout 400 python tools/e2e_chunk_probe.py > gpurun_out/r3u/chunks.jsonl 2> gpurun_out/r3u/chunks.err; cat gpurun_out/r3u/chunks.jsonl; tail -2 gpurun_out/r3u/chunks.err
