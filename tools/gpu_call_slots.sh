#!/bin/bash
# decoder scratch by compressed-block slots: smoke + every test that decodes Zstandard on the device
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python -m pytest tests/test_gpu_zstd_dec.py tests/test_gpu_zstd_long.py tests/test_gpu_zstd_enc.py tests/test_7z_writer.py tests/test_ref_7z_host.py -m gpu -x -q 2>&1 | tail -4
