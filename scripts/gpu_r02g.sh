#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -u -m pytest tests/test_gpu_parity.py -m gpu --tb=short --timeout 180 -p no:cacheprovider -q -s -k "match_pairs or match_sharded" 2>&1 | tail -8
echo "== frames config"; timeout 600 python bench.py --frames 64 --steps 3 --warmup 3 --noise-frames 32 > gpurun_out/r02g_ours.json 2> gpurun_out/r02g_ours.err; tail -c 2500 gpurun_out/r02g_ours.json; tail -3 gpurun_out/r02g_ours.err
echo "== stream config"; timeout 600 python bench.py --config stream --stream-frames 16 --steps 3 --warmup 3 > gpurun_out/r02g_stream.json 2> gpurun_out/r02g_stream.err; tail -c 2500 gpurun_out/r02g_stream.json; tail -3 gpurun_out/r02g_stream.err
echo "== reference, frames"; timeout 600 python bench.py --impl reference --frames 32 --steps 2 --warmup 1 --noise-frames 8 > gpurun_out/r02g_ref.json 2> gpurun_out/r02g_ref.err; tail -c 1500 gpurun_out/r02g_ref.json; tail -3 gpurun_out/r02g_ref.err
echo "== reference, stream"; timeout 600 python bench.py --impl reference --config stream --stream-frames 4 --steps 2 --warmup 1 > gpurun_out/r02g_refstream.json 2> gpurun_out/r02g_refstream.err; tail -c 1500 gpurun_out/r02g_refstream.json; tail -3 gpurun_out/r02g_refstream.err
