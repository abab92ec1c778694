APAP_B200_LIB=cvx_proj_b200/lab/trace.so python tools/gram_scan.py 3776:5120 | head -50
