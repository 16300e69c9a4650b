export FMI_ATTN_PAIR=0
timeout 300 python -m pytest tests/test_attention_gpu.py -m gpu -q -x --timeout 60 2>&1 | tail -3
echo "--- v2 wide QK (N=128)"; timeout 500 python tools/perf/perf_attention.py 2>&1 | grep "128x128 N=8\|64x64 N=32\|32x32"
echo "--- v2 64-key QK"; FMI_ATTN_DBG=32 timeout 500 python tools/perf/perf_attention.py 2>&1 | grep "128x128 N=8\|64x64 N=32"
