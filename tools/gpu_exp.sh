#!/bin/bash
for e in 1 2 3; do echo "experiment $e"; B200Z_LIB=tools/mb/libb200z_exp$e.so timeout 60 python tools/tools_profile_enc.py 4096 20 2 2>&1 | tail -1; done
