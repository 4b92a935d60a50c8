timeout 120 python tools/sat_microbench.py 2>&1 | tail -1
DSTAG_SAT_MMA=0 timeout 120 python tools/sat_microbench.py 2>&1 | tail -1
timeout 60 python tools/sat_microbench.py --N 307 --B 32 2>&1 | tail -1
timeout 60 python tools/sat_microbench.py --N 1021 --B 8 --K 2 2>&1 | tail -1
timeout 60 python tools/sat_microbench.py --N 9 --B 2 2>&1 | tail -1
timeout 60 python tools/sat_microbench.py --N 170 --B 32 2>&1 | tail -1
timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -k spatial 2>&1 | tail -3
