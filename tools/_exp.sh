DGS_BLOCKS_TRACE=1 timeout 60 python tools/timing_run.py 2>&1 | tail -7
