RG_APPLY_VARIANT_TEST=2 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k zslab 2>&1 | grep -E "Error|error|assert|Mismatch|Max|slab|x:|y:" | head -30
timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -p no:cacheprovider 2>&1 | tail -5
